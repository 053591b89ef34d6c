"""e2e verify (pinned host buffers) for different sub-chunk schedules (S256_VERIFY_CUTS, in 64ths)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cuts in (os.environ.get("CUTS_LIST", "4,16;8").split(";")):
    env = dict(os.environ, S256_VERIFY_CUTS=cuts)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "e2e_trace.py")], env=env, capture_output=True, text=True, timeout=600)
    print(cuts, (p.stdout.strip().splitlines() or [p.stderr[-300:]])[-1], flush=True)
