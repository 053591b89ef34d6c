"""F_p multiplication probes: dependent products per second for the ladders'
IMAD.WIDE multiplier, out of line and inlined."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1024)
names = ["fe_mul 8x32 IMAD.WIDE, out of line (as in k_dsm)", "fe_mul 8x32 IMAD.WIDE, inlined"]
out = {}
for form, nm in enumerate(names):
    best = max(eng.microbench_fe_mul(form, 2048)[0] for _ in range(3))
    out[nm] = {"muls_per_s": best, "clk_sm_per_mul_at_1965MHz": 148 * 1.965e9 / best}
    print(nm, f"{best/1e9:.1f} G mul/s  ({148*1.965e9/best:.3f} SM-clocks per product)")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench_fe_mul.json", "w"), indent=1)
