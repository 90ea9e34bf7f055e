"""Same-box A/B of the mapping-network kernels across library builds: loads each given .so with a minimal ctypes binding (only the
entry points every build has) and times gsb_mapping_forward on 1.01M rows.  usage: python tools/ab_mapping.py lib1.so lib2.so ..."""
import ctypes as C
import sys
import numpy as np
import torch

rows = 1_010_000
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
w = torch.tensor((rng.standard_normal((8, 512, 512)) * 100).astype(np.float32), device=dev)      # EqualLinear init scale (randn / lr_mul)
b = torch.zeros((8, 512), dtype=torch.float32, device=dev)
z = torch.randn((rows, 512), device=dev, dtype=torch.float32)
out = torch.empty_like(z)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rnd in range(2):
    for path in sys.argv[1:]:
        lib = C.CDLL(path)
        lib.gsb_mapping_packed_bytes.restype = C.c_size_t
        lib.gsb_mapping_packed_bytes.argtypes = [C.c_int, C.c_int]
        lib.gsb_mapping_workspace_bytes.restype = C.c_size_t
        lib.gsb_mapping_workspace_bytes.argtypes = [C.c_int64, C.c_int]
        lib.gsb_mapping_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        lib.gsb_mapping_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gsb_last_error.restype = C.c_char_p
        packed = torch.empty(lib.gsb_mapping_packed_bytes(8, 512), dtype=torch.uint8, device=dev)
        ws = torch.empty(lib.gsb_mapping_workspace_bytes(rows, 512), dtype=torch.uint8, device=dev)
        assert lib.gsb_mapping_pack(w.data_ptr(), b.data_ptr(), 8, 512, C.c_float(0.01), packed.data_ptr(), st) == 0

        def fwd():
            rc = lib.gsb_mapping_forward(packed.data_ptr(), 8, 512, z.data_ptr(), out.data_ptr(), rows, 1, ws.data_ptr(), ws.numel(), st)
            assert rc == 0, lib.gsb_last_error()
        for _ in range(2):
            fwd()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fwd(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"round {rnd} {path}: ms {[round(t, 2) for t in ts]} checksum {float(out[:1000].double().sum()):.6f}", flush=True)
        del packed, ws
