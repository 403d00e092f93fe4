"""Per-kernel timing of the four north-star kernels (K1 z-step, K2 Phi draw, K3 count rebuild, K4
log-likelihood) on the PubMed-shaped shard, through the step-wise C ABI.

    python tools/kernel_table.py [--workload pubmed8] [--reps 5] > gpurun_out/kernel_table.json

Each phase is a blocking library call (kernel launches + one stream synchronise); the wall time of the
call is the device time plus ~10 us of launch/sync overhead, which is negligible for these ms-scale
phases.  Run the same command under `ncu --set full` for the DRAM bytes (profiles/README.md).
Algorithmic bytes: SURVEY.md 8(d).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import ldagroupedgibbssampler_b200 as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pubmed8")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--docs", type=int, default=0)
    a = ap.parse_args()
    wl = dict(B.WORKLOADS[a.workload])
    if a.docs:
        wl["D"] = a.docs
    off, tokens = L.synth_corpus(wl["D"], wl["V"], wl["mean_len"], seed=B.CORPUS_SEED)
    N, D, K, V = len(tokens), wl["D"], wl["K"], wl["V"]
    cfg = L.LDAConfiguration(scheme=wl["scheme"], topics=K, alpha=wl["alpha"], beta=wl["beta"], seed=B.SEED, exec_time=0)
    s = L.GpuLDASampler(cfg, device=0)
    s.addInstances(L.InstanceList.from_csr(off, tokens, V))
    s.sample(2)                                   # past the random initial state
    peak, src = B.peaks()
    ggs = wl["scheme"] == "gpu_ggs"

    def timed(fn):
        fn()
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(ts))

    rows = []

    def row(name, kernels, ms, alg_bytes, note=""):
        gbs = alg_bytes / (ms / 1e3) / 1e9
        # SURVEY 8(d) algorithmic bytes over the phase's time: a MODEL figure (for the z-step it charges one Phi^T row per
        # token and exceeds the HBM peak when rows are shared or cached); the roofline fraction from measured DRAM bytes
        # is in profiles/z_kernel_traffic.json and in bench.py's `roofline`
        rows.append({"phase": name, "kernels": kernels, "ms": round(ms, 4), "algorithmic_bytes": int(alg_bytes),
                     "model_gbs": round(gbs, 1), "frac_model": round(gbs / peak, 4), "note": note})

    if ggs:
        row("theta draw (GGS)", "theta_kernel", timed(lambda: s._step("sample_theta")), 4 * N + 4 * K * D,
            "reads z, writes theta [D][K] fp32; instruction bound (K*D Gamma draws)")
    row("K1 z-step", "z_kernel", timed(lambda: s._step("sample_z")), (4 * K + 12) * N,
        "one fp32 K-vector of Phi^T per token + w + z in + z out; rows shared by runs and served by L2")
    row("K3 count rebuild (stand-alone)", "counts_kernel + topic_totals_kernel", timed(lambda: s._step("rebuild_counts")),
        16 * N + 4 * K * V, "setZIndicators / ldagpu_rebuild_counts path; inside a sweep the counts are fused into K1")
    row("K2 Phi draw", "phi_draw + phi_segment + phi_normalise", timed(lambda: s._step("sample_phi")), 8 * K * V,
        "reads n_wk, writes Phi^T; fp64 Gamma draws: instruction bound")
    row("K4 log-likelihood", "ll_doc + ll_type + sum_partials", timed(lambda: s.modelLogLikelihood()),
        4 * N + 4 * K * V + 8 * D, "includes the D2H of n_k and the host-side parameter terms")
    row("log-posterior", "lp_tokens + lp_theta + lp_phi", timed(lambda: s.computeLogPosterior()),
        12 * N + 4 * N + 4 * K * D + 4 * K * V, "reads w, z, one Phi entry per token (32 B sector), theta, Phi")
    sweep_ms = timed(lambda: s.sample(1))
    print(json.dumps({"workload": wl["desc"], "N": N, "D": D, "K": K, "V": V, "peak_gbs": peak, "peak_source": src,
                      "sweep_ms": round(sweep_ms, 3), "rows": rows}, indent=1))
    s.close()


if __name__ == "__main__":
    main()
