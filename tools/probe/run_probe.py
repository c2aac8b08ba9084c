import ctypes, os, sys, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "libprobe.so"))
lib.probe_umma.argtypes = [ctypes.c_void_p] * 1 + [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3
torch.manual_seed(0)
rows = 360
buf = (torch.randint(-8, 9, (rows, 32)).float() / 8).cuda()      # exactly representable in tf32
B = (torch.randint(-8, 9, (64, 32)).float() / 8).cuda()
for sbo_rows in (8, 10, 16, 18):
    for row_off in (0, 1, 3, 11):
        for mode in (0, 1):
            D = torch.zeros(128, 64, device="cuda")
            rc = lib.probe_umma(buf.data_ptr(), rows, B.data_ptr(), D.data_ptr(), row_off, sbo_rows, mode)
            idx = torch.tensor([row_off + (m // 8) * sbo_rows + (m % 8) for m in range(128)], device="cuda")
            ref = buf[idx] @ B.t()
            err = float((D - ref).abs().max())
            print(f"sbo_rows={sbo_rows:2d} row_off={row_off:2d} base_offset_mode={mode} rc={rc} max_abs_err={err:.3g} {'OK' if err == 0 else 'MISMATCH'}", flush=True)
