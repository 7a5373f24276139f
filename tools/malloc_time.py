"""cudaMalloc / cudaFree / cudaMemset cost by size (why the POA context pre-sizes its trace buffer)."""
import ctypes as C, time, glob, sys
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = C.CDLL(glob.glob(torch.__file__.rsplit("/", 1)[0] + "/../nvidia/cuda_runtime/lib/libcudart.so*")[0])
for gb in (0.25, 1, 4, 8, 8, 16):
    n = int(gb * (1 << 30)); p = C.c_void_p()
    t0 = time.perf_counter(); rc = rt.cudaMalloc(C.byref(p), C.c_size_t(n)); t1 = time.perf_counter()
    rt.cudaMemset(p, 0, C.c_size_t(n)); rt.cudaDeviceSynchronize(); t2 = time.perf_counter()
    rt.cudaFree(p); t3 = time.perf_counter()
    print("%.2f GB rc=%d malloc %.1f ms memset %.1f ms free %.1f ms" % (gb, rc, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
