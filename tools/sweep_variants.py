"""Sweep the solve-sweep kernel variants (WS_FWD_KB / WS_BWD_CB / WS_BWD_RJ [+ _LEAF]) on the 1M-element
vector problem; diagnostic only."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
combos = [dict(), dict(WS_FWD_KB="2"), dict(WS_FWD_KB="8"), dict(WS_FWD_KB_LEAF="4"), dict(WS_FWD_KB_LEAF="2"),
          dict(WS_BWD_CB="1"), dict(WS_BWD_CB="4"), dict(WS_BWD_CB_LEAF="1"), dict(WS_BWD_CB_LEAF="4"),
          dict(WS_BWD_RJ="2"), dict(WS_BWD_RJ_LEAF="2"),
          dict(WS_FWD_KB="2", WS_FWD_KB_LEAF="4", WS_BWD_CB="1", WS_BWD_CB_LEAF="1", WS_BWD_RJ="2", WS_BWD_RJ_LEAF="2")]
if len(sys.argv) > 1:
    combos = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
for c in combos:
    env = dict(os.environ, **c)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "solve_only.py"), "--solves", "30"], env=env,
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    print(c, out, flush=True)
