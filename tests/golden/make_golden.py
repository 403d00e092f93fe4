"""Regenerates the fixtures under tests/golden/.  Run in the BUILD container only (it reads
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

cats_corpus.npz   the reference's bundled corpus src/main/resources/datasets/cats.txt
                  (D=23, V=303, N=7788) tokenised as LDAUtils.loadInstancesPrune does with the
                  settings of configuration/plda-cats-test.cfg (empty stop list, rare_threshold=0,
                  keep_numbers): lower-case, first-seen vocabulary order.  Type ids only -- derived
                  data, no reference source code.
oracle_golden.npz outputs of the CPU oracle (oracle/lda_oracle.c) on that corpus with the run
                  configuration of BASELINE.json configs[0]: K=20 (--topics=20), alpha=5, beta=7,
                  seed 2019; contract AND faithful mode, GGS and PCGS, 3 sweeps.  These pin the
                  oracle against silent changes; they are NOT outputs of the Java reference (no JDK
                  here -- "parity unpinned", see oracle/lda_oracle.h).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ldagroupedgibbssampler_b200.corpus import load_dataset  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CATS = "/root/reference/src/main/resources/datasets/cats.txt"


def main():
    il = load_dataset(CATS)
    off, tokens = il.to_csr()
    assert (il.size(), il.alphabet.size(), len(tokens)) == (23, 303, 7788), (il.size(), il.alphabet.size(), len(tokens))
    np.savez_compressed(os.path.join(HERE, "cats_corpus.npz"), doc_offsets=off, tokens=tokens)

    K, V, alpha, beta, seed = 20, 303, 5.0, 7.0, 2019
    al = np.full(K, alpha)
    z0 = O.java_next_ints(seed, K, len(tokens))
    n_wk0, n_k0 = O.rebuild_counts(tokens, z0, V, K)
    out = dict(z0=z0, n_k0=n_k0)
    for mode in ("contract", "faithful"):
        phi0 = O.phi_contract(n_wk0, beta, seed, 0) if mode == "contract" else O.phi_faithful(n_wk0, beta, seed, 0)
        for scheme, name in ((O.GGS, "ggs"), (O.PCGS, "pcgs")):
            st = dict(z=z0, phiT=phi0)
            lls = []
            for it in (1, 2, 3):
                st = O.sweeps(mode, scheme, off, tokens, st["z"], V, K, al, beta, seed, it, 1, st["phiT"])
                lls.append(O.log_likelihood(off, st["z"], K, V, st["n_wk"], st["n_k"], al, beta))
            out[f"{mode}_{name}_z3"] = st["z"]
            out[f"{mode}_{name}_nk3"] = st["n_k"]
            out[f"{mode}_{name}_ll"] = np.array(lls)
            out[f"{mode}_{name}_phi3_colsum"] = st["phiT"].astype(np.float64).sum(axis=0)
            out[f"{mode}_{name}_phi3_sample"] = st["phiT"][::37, ::3].copy()
            if scheme == O.GGS:
                out[f"{mode}_{name}_theta3_sample"] = st["theta"][::5, ::3].copy()
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
