"""Probe: achievable INT8 tensor throughput on this B200 (cuBLASLt through torch._int_mm) -- the
ceiling an Ozaki-scheme FP64 emulation could draw on."""
import torch, time
for n in (8192, 13824):
    a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    bt = b.t().contiguous().t()   # column-major B
    for name, B in (("rowmajorB", b), ("colmajorB", bt)):
        try:
            torch._int_mm(a, B); torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch._int_mm(a, B); e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"int8 {n}^3 {name}: {2.0*n**3/(best*1e-3)/1e12:.0f} TOPS ({best:.2f} ms)", flush=True)
        except Exception as e:
            print("int8", n, name, "failed:", str(e)[:200])
    x = torch.randn(n, n, dtype=torch.bfloat16, device="cuda"); y = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
    torch.matmul(x, y); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(x, y); e1.record(); e1.synchronize()
    print(f"bf16 {n}^3: {2.0*n**3/(e0.elapsed_time(e1)*1e-3)/1e12:.0f} TFLOPS", flush=True)
    x8 = x.to(torch.float8_e4m3fn); y8 = y.t().contiguous().to(torch.float8_e4m3fn).t()
    try:
        s = torch.tensor(1.0, device="cuda")
        torch._scaled_mm(x8, y8, scale_a=s, scale_b=s, out_dtype=torch.bfloat16); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch._scaled_mm(x8, y8, scale_a=s, scale_b=s, out_dtype=torch.bfloat16); e1.record(); e1.synchronize()
        print(f"fp8 {n}^3: {2.0*n**3/(e0.elapsed_time(e1)*1e-3)/1e12:.0f} TFLOPS", flush=True)
    except Exception as e:
        print("fp8 failed:", str(e)[:200])
