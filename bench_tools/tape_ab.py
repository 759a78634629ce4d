"""Developer tool (GPU box): tape scheduling (GF_TAPE=1) vs the uniform band split (default).
Reports the difference of the two outputs (last-bit: re-seed phases differ) and times both on gray / colour / batch / giga cases."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)


def setenv(env):
    for k, v in env.items(): api.set_option(k, int(v))


def clrenv(env):
    for k in env: api.set_option(k, -1)


def timed(f, iters):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): f()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gray(w, h, r, border=0, iters=40, nsets=4):
    g = torch.Generator(device="cuda").manual_seed(1)
    if (w * h * 12 * nsets) > 40e9: nsets = 1
    sets = [(torch.rand((h, w), device="cuda", generator=g), torch.rand((h, w), device="cuda", generator=g)) for _ in range(nsets)]
    q = [torch.full((h, w), float("nan"), device="cuda") for _ in range(2)]
    res = {"case": f"gray {w}x{h} r={r} border={border}"}
    for tag, env in (("tape", {"GF_TAPE": 1}), ("uniform", {"GF_TAPE": 0})):
        setenv(env)
        i = [0]
        def f():
            a, b = sets[i[0] % nsets]; i[0] += 1
            api.call("gf_guided_gray", a.data_ptr(), b.data_ptr(), q[tag == "tape"].data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, border, sp)
        ms = timed(f, iters)
        # one more call on set 0 for the comparison
        a, b = sets[0]
        api.call("gf_guided_gray", a.data_ptr(), b.data_ptr(), q[tag == "tape"].data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, border, sp)
        torch.cuda.synchronize()
        res[tag + "_us"] = round(ms * 1e3, 2); res[tag + "_kernel"] = api.last_kernel()
        clrenv(env)
    res["max_diff"] = float((q[0] - q[1]).abs().max()); res["nan"] = int(torch.isnan(q[1]).sum())
    res["gain_pct"] = round(100 * (1 - res["tape_us"] / res["uniform_us"]), 1)
    print(json.dumps(res), flush=True)


def color(n, w, h, r, iters=5, extra=None):
    g = torch.Generator(device="cuda").manual_seed(2)
    I = torch.rand((n, h, w, 3), device="cuda", generator=g); p = torch.rand((n, h, w), device="cuda", generator=g)
    q = [torch.full((n, h, w), float("nan"), device="cuda") for _ in range(2)]
    res = {"case": f"colour {n}x{w}x{h} r={r}", "extra": extra or {}}
    for tag, env in (("tape", dict(extra or {}, GF_TAPE=1)), ("uniform", {"GF_TAPE": 0})):
        setenv(env)
        f = lambda: api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q[tag == "tape"].data_ptr(), n, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, sp)
        ms = timed(f, iters)
        res[tag + "_ms"] = round(ms, 4); res[tag + "_kernel"] = api.last_kernel()
        clrenv(env)
    res["max_diff"] = float((q[0] - q[1]).abs().max()); res["nan"] = int(torch.isnan(q[1]).sum())
    res["gain_pct"] = round(100 * (1 - res["tape_ms"] / res["uniform_ms"]), 1)
    print(json.dumps(res), flush=True)
    del I, p, q
    torch.cuda.empty_cache()


cases = sys.argv[1:] or ["gray", "color", "giga"]
if "gray" in cases:
    for (w, h, r, b) in [(3840, 2160, 8, 0), (3840, 2160, 8, 0), (7680, 4320, 8, 0), (1920, 1080, 8, 0), (3840, 2160, 4, 0), (3840, 2160, 16, 0),
                         (7680, 4320, 16, 0), (7680, 4320, 32, 0), (3840, 2160, 8, 1), (3840, 2160, 7, 2), (1000, 700, 5, 0), (16384, 8192, 8, 0)]:
        gray(w, h, r, b)
if "color" in cases:
    color(32, 1920, 1080, 16)
    color(32, 1920, 1080, 16, extra={"GF_C4_EDGE_WEIGHT": 110})
    color(32, 1920, 1080, 16, extra={"GF_C4_EDGE_WEIGHT": 120})
    color(1, 1920, 1080, 16, iters=20)
    color(8, 1920, 1080, 8)
    color(3, 1280, 720, 4, iters=20)
    color(256, 1920, 1080, 16, iters=3)
if "giga" in cases:
    gray(32768, 32768, 16, iters=5, nsets=1)
