"""scripts/lncde_sweep.py end to end without a GPU (two tiny points over the host-emulated kernel library); run by
tests/test_bench_dryrun_cpu.py in a subprocess.  Guards the script's own Python; its numbers are meaningless."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, REPO, os.path.join(REPO, "log-neural-cdes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gpu_on_emu_plugin

gpu_on_emu_plugin.pytest_configure(None)
sys.argv = ["lncde_sweep.py", "--no-graph", "--steps", "1", "--warmup", "1", "--grid", "2,9,16,3;2,70,8,2"] + sys.argv[1:]
runpy.run_path(os.path.join(REPO, "scripts", "lncde_sweep.py"), run_name="__main__")
