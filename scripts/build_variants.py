"""Builds tuning variants of the library into secp256k1-voi_b200/lib/variants/ (local, no GPU)."""
import importlib, os, re, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bld = importlib.import_module("secp256k1-voi_b200.build")
VARIANTS = dict(a.split("=", 1) for a in sys.argv[1:])  # name="-DX=1 -DY=2"
outdir = os.path.join(bld.LIBDIR, "variants")

def one(item):
    name, flags = item
    out = os.path.join(outdir, name + ".so")
    bld.build(extra=flags.split(), out=out)
    log = open(out + ".log").read()
    m = re.search(r"Compiling entry function '_Z5k_dsm.*?Used (\d+) registers", log, re.S)
    sp = re.search(r"Compiling entry function '_Z5k_dsm.*?(\d+) bytes spill stores", log, re.S)
    return name, m.group(1) if m else "?", sp.group(1) if sp else "?"

with ThreadPoolExecutor(4) as ex:
    for name, regs, spill in ex.map(one, VARIANTS.items()):
        print(f"{name}: k_dsm regs={regs} spill_stores={spill}")
