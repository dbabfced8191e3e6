"""Extended randomized parity sweep: runs tests/test_gpu_parity.py::test_randomized_parity's body for many
more seeds than the test suite does (python bench_micro/fuzz_parity.py FIRST LAST)."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import test_gpu_parity as T
a, b = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(a, b):
    try:
        T.test_randomized_parity.__wrapped__(seed) if hasattr(T.test_randomized_parity, "__wrapped__") else T.test_randomized_parity(seed)
    except Exception as e:          # noqa: BLE001
        bad.append((seed, repr(e)[:200]))
        print("seed", seed, "FAILED", repr(e)[:200], flush=True)
print("seeds %d..%d: %d failures" % (a, b - 1, len(bad)))
sys.exit(1 if bad else 0)
