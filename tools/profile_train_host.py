"""Diagnostics (GPU): cProfile of the host side of the config-3 training step (tools/bench_train_chunk.py)."""
import cProfile, pstats, sys, io, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.argv = ["x", "0", "20"]
import tools.bench_train_chunk as T
pr = cProfile.Profile()
pr.enable()
T.main()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
