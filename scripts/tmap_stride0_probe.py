"""Does cuTensorMapEncodeTiled accept zero strides (a virtual nearest-neighbour upsampling view of a tensor)?"""
import ctypes
import torch

torch.cuda.init()
x = torch.zeros(4, 16, 16, 64, dtype=torch.bfloat16, device="cuda")  # (N, H/2, W/2, C)
cu = ctypes.CDLL("libcuda.so.1")
n, h2, w2, c = x.shape
for name, dims, strides in [
    ("plain 4d", [c, w2, h2, n], [c * 2, c * 2 * w2, c * 2 * w2 * h2]),
    ("upsampled 5d (zero strides)", [c, 2, w2, 2, n * h2], [0, c * 2, 0, c * 2 * w2]),
    ("upsampled 5d (stride 16 B dummy)", [c, 2, w2, 2, n * h2], [16, c * 2, 16, c * 2 * w2]),
]:
    rank = len(dims)
    tm = (ctypes.c_uint64 * 16)()
    gd = (ctypes.c_uint64 * rank)(*dims)
    gs = (ctypes.c_uint64 * (rank - 1))(*strides)
    box = (ctypes.c_uint32 * rank)(*([64, 2, 6, 2, 10][:rank] if rank == 5 else [64, 10, 18, 1]))
    es = (ctypes.c_uint32 * rank)(*([1] * rank))
    rc = cu.cuTensorMapEncodeTiled(tm, 9, rank, ctypes.c_void_p(x.data_ptr()), gd, gs, box, es, 0, 3, 3, 0)
    print(f"{name}: rc = {rc}")
