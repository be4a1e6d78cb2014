"""Run bench.py after calling validation hooks of the library: `python scripts/bench_with.py cta_group=1 -- --steps 20 ...`
(hooks: cta_group, group_m, decode_merged; see the end of include/ospo_head.h)."""
import runpy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ospo_b200 import _abi  # noqa: E402

i = sys.argv.index("--")
lib = _abi.load()
for kv in sys.argv[1:i]:
    k, v = kv.split("=")
    getattr(lib, f"ospo_head_set_{k}")(int(v))
sys.argv = [str(ROOT / "bench.py")] + sys.argv[i + 1:]
runpy.run_path(str(ROOT / "bench.py"), run_name="__main__")
