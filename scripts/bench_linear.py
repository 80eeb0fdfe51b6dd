"""Times the tcgen05 split-bf16 linear alone through the C ABI (meshtok_linear) at the encoder's shapes."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr

M = 51200
shapes = [(192, 192), (64, 384), (64, 64), (128, 128), (256, 256), (256, 512), (576, 512), (576, 576)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for N, K in shapes:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    nb = C.c_size_t()
    check(lib.meshtok_linear_image_bytes(N, K, C.byref(nb))); wimg = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_linear_build_image(ptr(W), N, K, ptr(wimg), stream_ptr()))
    check(lib.meshtok_rows_image_bytes(M, K, C.byref(nb))); aimg = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_rows_build_image(ptr(A), M, K, ptr(aimg), stream_ptr()))
    out = torch.empty(M, N, device="cuda")
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.meshtok_linear(ptr(A), ptr(aimg), ptr(W), ptr(wimg), ptr(b), ptr(out), M, N, K, 0, 0, stream_ptr()))
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"N={N:4d} K={K:4d}: {t:7.1f} us  {2*M*N*K*3/t/1e6:7.0f} TF/s on the tensor pipe (cold L2)")
