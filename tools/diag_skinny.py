import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
cabi = sys.modules[pkg.__name__ + "._cabi"]; L = cabi.load_library()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for cl in (1,):
    os.environ["FMT_SK_CLUSTER"] = str(cl)
    for M in (16, 48):
        N, K = 512, 1024
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        out = torch.full((M, N), float("nan"), device="cuda")
        rc = L.fmt_debug_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, -1, st)
        torch.cuda.synchronize()
        ref = (A.double() @ W.double().t()).float()
        bad = ((out - ref).abs() > 1e-2)
        rows = bad.any(1).nonzero().flatten().tolist()
        r0 = rows[0] if rows else 0
        print("out", out[r0, :6].tolist()); print("ref", ref[r0, :6].tolist())
        # does out row r0 equal some other ref row / partial-K sum?
        for kk in (64, 128, 256, 512):
            part = (A[r0, :kk].double() @ W[:6, :kk].double().t()).float()
            print("  partial K", kk, part.tolist())
        d = (ref[:, :6] - out[r0, :6]).abs().sum(1); print("  closest ref row", int(d.argmin()), float(d.min()))
        print(f"CL={cl} M={M} rc={rc} bad_rows={len(rows)} first={rows[:6]} last={rows[-3:]} badcols={bad.any(0).sum().item()}")
