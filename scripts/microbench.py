import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("secp256k1-voi_b200")
eng = pkg.Engine(device=0, max_batch=1024)
names = ["mad_wide(IMAD.WIDE.U32)", "madc_chain(IMAD.WIDE.U32.X)", "mad_lo(IMAD)", "mad_hi(IMAD.HI)", "addc_chain(IADD3.X)", "madc_pair(carry-out only)+addc", "mix(IMAD.WIDE + 2 IADD3.X each)",
         "dfma(DFMA rz)", "dfma+madc_chain 2:1 (ops = DFMA + IMAD.WIDE)", "dfma_prod(52x52 product: 2 DFMA + DADD + 2 x 64-bit add; +1 DADD feeding the next)",
         "dfma_prod+madc 2:1 (ops = products + IMAD.WIDE)"]
out = {}
for v, nm in enumerate(names):
    best = 0
    for _ in range(3):
        r, ms = eng.microbench_variant(v, 4096)
        best = max(best, r)
    out[nm] = {"ops_per_s": best, "per_clk_per_sm_at_1965MHz": best / 148 / 1.965e9}
    print(nm, f"{best/1e12:.3f} Tops/s  ({best/148/1.965e9:.1f} /clk/SM @1965MHz)")
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
