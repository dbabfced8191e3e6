"""Build bench_micro/libctb_exp.so: the library with -DCTB_EXPERIMENT (the streaming kernel then
honours CTB_KNOBS / CTB_STAGES / CTB_CHUNK_TB).  Use it with CTB_LIBRARY=bench_micro/libctb_exp.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

OUT = os.path.join(ROOT, "bench_micro", "libctb_exp.so")


def build(out=OUT, defines=()):
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + G.NVCC_FLAGS + \
          ["-DCTB_EXPERIMENT"] + ["-D" + d for d in defines] + \
          ["-I" + os.path.join(ROOT, "include"), "-shared", "-o", out] + \
          [os.path.join(G.CSRC, s) for s in G.SOURCES]
    subprocess.run(cmd, check=True, cwd=G.CSRC)
    return out


if __name__ == "__main__":
    # python bench_micro/build_exp.py [suffix DEF=VAL ...]
    if len(sys.argv) > 1:
        print(build(OUT.replace(".so", "_" + sys.argv[1] + ".so"), sys.argv[2:]))
    else:
        print(build())
