"""Sweep the assembly-kernel launch knobs (WS_ASM_RANGES / WS_ASM_PREFETCH / WS_ASM_MINB) on the
1M-element problems; diagnostic only."""
import os, subprocess, sys, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for minb in (4, 6, 8):
    for ranges in (1, 2, 16):
        for pf in (0, 1):
            env = dict(os.environ, WS_ASM_RANGES=str(ranges), WS_ASM_PREFETCH=str(pf), WS_ASM_MINB=str(minb))
            out = subprocess.run([sys.executable, os.path.join(root, "tools", "microbench.py"), "--solver", "0"], env=env,
                                 capture_output=True, text=True).stdout.strip().splitlines()[-1]
            d = json.loads(out)
            print("minb %d ranges %2d prefetch %d: scalar %.3f ms  vector %.3f ms  qd %.3f ms" % (minb, ranges, pf, d["scalar_assemble_ms"], d["vector_assemble_ms"], d["vector_qd_ms"]), flush=True)
