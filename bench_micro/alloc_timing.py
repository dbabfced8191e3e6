import time, torch
dev = torch.device("cuda", 0)
main = torch.cuda.current_stream(dev)
def tm(f):
    t = time.perf_counter(); r = f(); return r, (time.perf_counter() - t) * 1e3
keep = None
for it in range(8):
    out, t_out = tm(lambda: torch.empty((1, 24378, 1460), dtype=torch.float64, device=dev))
    b0, t_b0 = tm(lambda: torch.empty((72, 349712), dtype=torch.float32, device=dev)); b0.record_stream(main)
    b1, t_b1 = tm(lambda: torch.empty((72, 349712), dtype=torch.float32, device=dev)); b1.record_stream(main)
    out.zero_()
    host, t_h = tm(lambda: torch.empty(out.shape, dtype=out.dtype, pin_memory=True))
    host.copy_(out, non_blocking=True)
    main.synchronize()
    keep = host.numpy()      # previous result dies here
    del out, b0, b1, host
    print("iter %d: out %.2f  dbuf %.2f %.2f  pinned %.2f ms" % (it, t_out, t_b0, t_b1, t_h), flush=True)
