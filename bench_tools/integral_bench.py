"""Developer tool: integral image (uint8 -> int32 / int64 SAT) timings, reduce-then-scan form against the round-1 two-pass
form (GF_SAT_TWO_PASS=1); the reference reports 0.597 ms at 4K on sm_86.

    python bench_tools/integral_bench.py > gpurun_out/integral.jsonl"""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for (w, h) in ((3840, 2160), (7680, 4320), (1920, 1080), (5910, 5941), (16384, 16384))[:int(os.environ.get("NCASES", "5"))]:
    for dt, name, fn in ((torch.int32, "i32", "gf_integral_u8_i32"), (torch.int64, "i64", "gf_integral_u8_i64")):
        nsets = 12 if w * h < 4e7 else 3
        sets = [(torch.randint(0, 256, (h, w), dtype=torch.uint8, device="cuda"), torch.empty((h, w), dtype=dt, device="cuda")) for _ in range(nsets)]
        scr = torch.empty((h // 16 + 2, w), dtype=dt, device="cuda")
        for form, opts in (("default", {}), ("reduce_then_scan", {"GF_SAT_TWO_PASS": 0}), ("two_pass_r1", {"GF_SAT_TWO_PASS": 1}), ("reduce_then_scan_hb32", {"GF_SAT_TWO_PASS": 0, "GF_SAT_HB": 32}),
                           ("reduce_then_scan_hb64", {"GF_SAT_TWO_PASS": 0, "GF_SAT_HB": 64})):
            for k, v in opts.items(): api.set_option(k, v)
            f = lambda i: api.call(fn, sets[i % nsets][0].data_ptr(), sets[i % nsets][1].data_ptr(), scr.data_ptr(), w, h, 0, 0, sp)
            for i in range(5): f(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 60 if w * h < 4e7 else 12
            e0.record(s)
            for i in range(n): f(i)
            e1.record(s); torch.cuda.synchronize()
            for k in opts: api.set_option(k, -1)
            us = e0.elapsed_time(e1) / n * 1e3
            ob = 4 if dt == torch.int32 else 8
            print(json.dumps({"w": w, "h": h, "out": name, "form": form, "us": round(us, 1), "gpix_s": round(w * h / us / 1e3, 1),
                              "alg_gb_s": round((1.0 + ob) * w * h / us / 1e3, 1), "frac_hbm_6535": round((1.0 + ob) * w * h / us / 1e3 / 6535.7, 3)}), flush=True)
        del sets, scr
        torch.cuda.empty_cache()
