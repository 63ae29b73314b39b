import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.build()
import parity_util as pu
r = pu.hifigan_step_errors(16, 32, steps=1)
print("SN_NATIVE", os.environ.get("XVA_SN_NATIVE", "1"), "loss", {k: "%.1e" % v for k, v in r["loss"][0].items()},
      "dgrad", {k: (v if not isinstance(v, float) else "%.2e" % v) for k, v in r["dgrad"].items() if not isinstance(v, dict)},
      "ggrad", {k: (v if not isinstance(v, float) else "%.2e" % v) for k, v in r["ggrad"].items() if not isinstance(v, dict)},
      "weights", {n: "%.2e" % r["weights"][n]["global"] for n in ("G", "mpd", "msd")})
