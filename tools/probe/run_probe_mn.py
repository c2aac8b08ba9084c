"""MN-major descriptor probe (see probe_umma_mn.cu): python tools/probe/run_probe_mn.py > gpurun_out/umma_mn_probe.txt"""
import ctypes, os, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "libprobe_mn.so"))
lib.probe_umma_mn.argtypes = [ctypes.c_void_p] * 1 + [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3
torch.manual_seed(0)
rows = 360
buf = (torch.randint(-8, 9, (rows, 32)).float() / 8).cuda()      # exactly representable in tf32
B = (torch.randint(-8, 9, (64, 32)).float() / 8).cuda()          # row = block*32 + pixel, 32 channels
Bt = B.view(2, 32, 32)                                            # [block][pixel][c]
for lbo_rows in (32, 18, 24, 33, 40):
    for row_off in (0, 1, 2, 3, 5, 8):
        for mode in (0, 1):
            D = torch.zeros(128, 64, device="cuda")
            rc = lib.probe_umma_mn(buf.data_ptr(), rows, B.data_ptr(), D.data_ptr(), row_off, lbo_rows, mode)
            ref = torch.zeros(128, 64, device="cuda")
            for b in range(4):
                a = buf[row_off + b * lbo_rows: row_off + b * lbo_rows + 32]        # [pixel][c]
                for nb in range(2):
                    ref[b * 32:(b + 1) * 32, nb * 32:(nb + 1) * 32] = a.t() @ Bt[nb]  # sum over pixels
            err = float((D - ref).abs().max())
            print(f"lbo_rows={lbo_rows:2d} row_off={row_off:2d} base_offset_mode={mode} rc={rc} max_abs_err={err:.3g} "
                  f"{'OK' if err == 0 else 'MISMATCH'}", flush=True)
