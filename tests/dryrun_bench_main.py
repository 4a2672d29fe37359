"""TEST INFRASTRUCTURE (tests/test_gpu_suite_dryrun.py): `python -m torch.distributed.run ... tests/dryrun_bench_main.py <bench args>`
runs bench.py's main on every rank with tests/dryrun_backend.py in the library's place -- the multi-rank control flow of the
bench (bootstrap, topology, collectives of the timing and of the host-buffer segment, rank-0 printing) without a GPU."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ["CHMY_DRYRUN"] = "1"
from helpers import install_dryrun_if_requested

install_dryrun_if_requested()
# extra bench arguments travel in the environment: torch.distributed.run's own parser trips over abbreviations like `--n`
sys.argv = [os.path.join(os.path.dirname(HERE), "bench.py")] + sys.argv[1:] + os.environ.get("DRYRUN_BENCH_ARGS", "").split()
runpy.run_path(sys.argv[0], run_name="__main__")
