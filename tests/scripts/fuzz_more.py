"""Extended fuzz run of tests/test_gpu_parity.py::test_fuzz_fused_vs_oracle with many more seeds (one-off confidence run)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import canny_edge_b200 as cb  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402
import test_gpu_parity as t  # noqa: E402

n0, n1 = int(sys.argv[1]), int(sys.argv[2])
ctx, oracle = cb.Context(0), Oracle()
for seed in range(n0, n1):
    t.test_fuzz_fused_vs_oracle(ctx, oracle, seed)
print(f"fuzz seeds {n0}..{n1 - 1}: {40 * (n1 - n0)} cases, all bit-exact")
