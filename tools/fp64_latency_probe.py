"""DFMA throughput against chains in flight per scheduler: the dependent-issue latency of the
FP64 pipe, which sizes how many independent chains a kernel must keep in flight."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from base_b200 import groundwork as gw  # noqa: E402

mhz = gw.device_info(0)["sm_clock_mhz"]
print(f"SM clock ceiling {mhz} MHz; 148 SMs x 4 schedulers; a warp DFMA occupies a scheduler's FP64 lanes 2 cycles")
for ctas in (1, 2, 4, 8):
    for ilp in (1, 2, 4, 8):
        r = gw.dfma_peak(0, ctas_per_sm=ctas, ilp=ilp, iters=1 << 14, warmup=2, reps=5)
        chains = ctas * 2 * ilp                       # per scheduler: 8 warps per CTA over 4 schedulers
        per_sched_per_clk = r["tflops"] * 1e12 / 2 / 32 / (148 * 4) / (mhz * 1e6)   # warp-DFMAs per cycle
        print(f"warps/sched {ctas * 2:2d} ilp {ilp}: chains/sched {chains:3d}  {r['tflops']:7.3f} TFLOP/s  "
              f"{per_sched_per_clk:.3f} warp-DFMA/clk/sched  => latency if latency-bound {chains / per_sched_per_clk:6.1f} clk")

print("\nWhat does a non-FP64 instruction cost next to a DFMA?  ilp 8, 16 warps/scheduler:")
base = None
for k, what in ((0, "nothing"), (1, "1 IMAD"), (2, "2 IMAD"), (-2, "2 ALU (add, xor)"), (-4, "4 ALU (add, xor)")):
    r = gw.dfma_peak(0, ctas_per_sm=8, ilp=8, int_per_fma=k, iters=1 << 14, warmup=2, reps=5)
    base = base or r["tflops"]
    print(f"  per DFMA {what:18s}: {r['tflops']:7.3f} TFLOP/s = {r['tflops'] / base:.3f} of the DFMA-only rate")
