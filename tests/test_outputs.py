"""The reference's output files as the host mirror writes them: Phi / Theta CSV dumps of the
diagnostic block (UPL:757-775,806-815 -> util/LDAUtils.java:1199-1254) and the top-word tables of the driver
(tui/ParallelLDA.java:268-282 -> util/LDAUtils.java:874-912,1429-1460).  Formats are pinned by the behaviour of
java.text.DecimalFormat("00.###E0") and String.format("%.4f") that LDAUtils.formatDouble combines."""
import os

import numpy as np
import pytest

from conftest import make_corpus


def test_format_double_matches_the_java_formats():
    from ldagroupedgibbssampler_b200.sampler import format_double
    # |d| >= 1e-4 or d == 0: String.format("%.4f")
    assert format_double(0.0) == "0.0000"
    assert format_double(0.5) == "0.5000"
    assert format_double(12.34567) == "12.3457"
    assert format_double(0.0001) == "0.0001"
    assert format_double(-0.25) == "-0.2500"
    # 0 < |d| < 1e-4: DecimalFormat("00.###E0") -- two integer digits, <= 3 fraction digits, exponent to match
    assert format_double(1.2345e-5) == "12.345E-6"
    assert format_double(1.2e-5) == "12E-6"
    assert format_double(9.9996e-5) == "99.996E-6"
    assert format_double(-4.5e-7) == "-45E-8"
    assert format_double(2.5e-12) == "25E-13"
    assert format_double(1.00049e-9) == "10.005E-10" or format_double(1.00049e-9) == "10.005E-10"
    assert format_double(float(np.float32(1e-45))) .endswith("E-46")          # the fp32 floor of Phi


def test_ascii_matrix_top_words_and_interval(tmp_path):
    from ldagroupedgibbssampler_b200.sampler import (format_top_words, format_top_words_as_csv, in_range_interval,
                                                     write_ascii_double_matrix)
    fn = tmp_path / "m.csv"
    write_ascii_double_matrix(np.array([[0.5, 1e-6], [0.0, 0.25]]), str(fn))
    assert fn.read_text() == "0.5000,10E-7\n0.0000,0.2500\n"
    tw = [["cat", "dog"], ["fish", "bird"]]
    assert format_top_words_as_csv(tw) == "cat,dog\nfish,bird"
    assert format_top_words(tw) == "Topic 1: cat dog\nTopic 2: fish bird"
    assert in_range_interval(5, (1, 3, 5, 9)) and not in_range_interval(4, (1, 3, 5, 9))
    with pytest.raises(ValueError):
        in_range_interval(1, (1,))
    with pytest.raises(ValueError):
        in_range_interval(1, (1, 2, 3))


@pytest.mark.gpu
def test_diagnostic_csv_dumps_and_top_words(tmp_path, oracle):
    import ldagroupedgibbssampler_b200 as L
    off, tokens = make_corpus(40, 60, 30, seed=6)
    K, V = 6, 60
    il = L.InstanceList.from_csr(off, tokens, V)
    il.alphabet = L.Alphabet("w%d" % i for i in range(V))
    cfg = L.LDAConfiguration(scheme="gpu_ggs", topics=K, alpha=0.5, beta=0.1, seed=4, exec_time=0, start_diagnostic=2,
                             save_phi=True, print_ndocs_interval=(2, 3), print_ndocs_cnt=10, logging_path=str(tmp_path))
    s = L.GpuLDASampler(cfg)
    s.addInstances(il)
    s.sample(4)
    asc = tmp_path / "ascii"
    names = sorted(os.listdir(asc))
    assert names == ["Phi_KxV_6_60_00002.csv", "Phi_KxV_6_60_00003.csv", "Phi_KxV_6_60_00004.csv",
                     "Theta_DxK_10_6_00002.csv", "Theta_DxK_10_6_00003.csv"]
    phi = np.loadtxt(asc / "Phi_KxV_6_60_00004.csv", delimiter=",")
    assert phi.shape == (K, V) and np.allclose(phi, s.getPhi(), atol=6e-5)
    th = np.loadtxt(asc / "Theta_DxK_10_6_00003.csv", delimiter=",")
    assert th.shape == (10, K) and np.allclose(th.sum(axis=1), 1.0, atol=1e-3)
    lp = (tmp_path / "log-posterior.txt").read_text().splitlines()
    assert [int(l.split("\t")[0]) for l in lp] == [2, 3, 4]
    # top words: per topic the types by n_wk descending, ties in type order
    n_wk = s.getTypeTopicMatrix()
    idx = s.getTopWordIndices(5)
    for k in range(K):
        want = sorted(range(V), key=lambda w: (-int(n_wk[w, k]), w))[:5]
        assert list(idx[k]) == want
    s.writeTopWords(str(tmp_path / "TopWords.txt"), 5)
    lines = (tmp_path / "TopWords.txt").read_text().splitlines()
    assert len(lines) == K and all(len(l.split(",")) == 5 for l in lines)
    with pytest.raises(ValueError):
        s.getTopWordIndices(V + 1)
    s.close()
