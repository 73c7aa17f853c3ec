"""Times the bf16 GEMM kernels in isolation (CUDA events, L2-warm, back to back) at the shapes of the B=32 step."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
cabi = sys.modules[pkg.__name__ + "._cabi"]; L = cabi.load_library()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = int(os.environ.get("GB_M", 5760))
shapes = [("qkv", M, 3072, 1024), ("proj", M, 1024, 1024), ("fc1", M, 4096, 1024), ("fc2", M, 1024, 4096)]
kinds = [int(x) for x in os.environ.get("GB_BN", "256,512").split(",")]
for name, m, n, k in shapes:
    A = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    W = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.zeros(n, device="cuda"); out = torch.empty(m, n, device="cuda")
    ref = None
    for bn in kinds:
        L.fmt_debug_gemm_bench(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), m, n, k, bn, 3, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        L.fmt_debug_gemm_bench(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), m, n, k, bn, 20, st)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{name:5s} M={m} N={n} K={k} kernel={bn:4d}: {us:7.1f} us  {2*m*n*k/us/1e6:7.1f} TFLOP/s")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): torch.matmul(A, W.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): torch.matmul(A, W.t())
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{name:5s} cuBLAS (torch.matmul bf16)          : {us:7.1f} us  {2*m*n*k/us/1e6:7.1f} TFLOP/s")
