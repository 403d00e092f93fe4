"""sample() + getZIndicators in one call on a GGS corpus of >= 8 Mi tokens: the z-step of the last sweep runs in 8
document-aligned parts and each part's indicators travel to the host under the z-step of the next (engine.cu
sweep_enqueue, `stream_out`).  The host buffer must hold exactly the sampler's z -- both widths, parts that start on
odd token offsets included -- and the sweep must still be the oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_streamed_z_read_back_ggs(oracle):
    import ldagroupedgibbssampler_b200 as L
    K, V, alpha, beta, seed = 16, 3000, 0.3, 0.05, 11
    off, tokens = L.synth_corpus(97000, V, 89.0, seed=21)
    N = len(tokens)
    assert N >= (8 << 20) + 50000
    cfg = L.LDAConfiguration(scheme="gpu_ggs", topics=K, alpha=alpha, beta=beta, seed=seed, exec_time=0)
    s = L.GpuLDASampler(cfg)
    s.addInstances(L.InstanceList.from_csr(off, tokens, V))
    z0 = s.get_z_flat()
    phi0 = s.getPhi().T.astype(np.float32).copy()
    out32 = np.full(N, -1, np.int32)
    s.sample(1, z_out=out32)
    st = oracle.sweeps("contract", oracle.GGS, off, tokens, z0, V, K, np.full(K, alpha), beta, seed, 1, 1, phi0)
    assert np.array_equal(out32, st["z"]) and np.array_equal(s.get_z_flat(), st["z"])
    assert np.array_equal(s.getTypeTopicMatrix(), st["n_wk"])
    out16 = np.full(N, 65535, np.uint16)
    s.sample(2, z_out=out16)                      # only the LAST sweep of the call streams
    assert np.array_equal(out16.astype(np.int32), s.get_z_flat())
    st2 = oracle.sweeps("contract", oracle.GGS, off, tokens, st["z"], V, K, np.full(K, alpha), beta, seed, 2, 2, st["phiT"])
    assert np.array_equal(out16.astype(np.int32), st2["z"])
    s.close()
