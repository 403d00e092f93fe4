"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute
without a GPU), configuration parsing, corpus containers, sharding."""
import ctypes as C
import os
import re
import textwrap

import numpy as np
import pytest

import ldagroupedgibbssampler_b200 as L
from conftest import ROOT, make_corpus


def test_library_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "ldagpu.h")).read()
    declared = set(re.findall(r"\b(ldagpu_[a-z0-9_]+)\s*\(", header))
    declared.discard("ldagpu_handle_s")
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.ldagpu_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    lib = L.load()
    if lib.ldagpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    off, tokens = make_corpus(5, 10, 4, seed=1)
    s = L.GpuLDASampler(L.LDAConfiguration(topics=3, alpha=0.1))
    with pytest.raises(L.LdaGpuError, match="no CUDA device"):
        s.addInstances(L.InstanceList.from_csr(off, tokens, 10))
    with pytest.raises(L.LdaGpuError):
        s.sample(1)


def test_create_argument_validation():
    lib = L.load()
    h = C.c_void_p()
    off = np.array([0, 2], np.int64)
    tok = np.array([0, 1], np.int32)
    al = np.array([0.1, 0.1], np.float64)

    def create(K=2, V=2, beta=0.01, scheme=0, alpha=al, offs=off):
        return lib.ldagpu_create(K, V, len(offs) - 1, L._lib.ptr(offs), L._lib.ptr(tok), L._lib.ptr(alpha), beta,
                                 1, scheme, 0, 0, 0, C.byref(h))

    assert create(K=0) != 0 and b"K" in lib.ldagpu_last_error(None)
    assert create(scheme=7) != 0 and b"scheme" in lib.ldagpu_last_error(None)
    assert create(beta=0.0) != 0 and b"beta" in lib.ldagpu_last_error(None)
    assert create(alpha=np.array([0.1, -1.0])) != 0 and b"alpha" in lib.ldagpu_last_error(None)
    assert create(offs=np.array([1, 2], np.int64)) != 0
    assert create(K=30000, alpha=np.full(30000, 0.1)) != 0 and b"too large" in lib.ldagpu_last_error(None)


def test_unknown_scheme_rejected():
    with pytest.raises(ValueError):
        L.GpuLDASampler(L.LDAConfiguration(scheme="ggs"))    # the Java schemes stay in Java
    assert set(L.SCHEMES) == {"gpu_ggs", "gpu_pcgs", "gpu_spalias", "gpu_polyaurn"}
    assert isinstance(L.createModel(L.LDAConfiguration(scheme="gpu_pcgs", topics=4)), L.GpuLDASampler)


def test_sampler_interface_is_complete():
    """Every method of the reference's sampler interfaces exists on the host mirror
    (topics/LDAGibbsSampler.java:10-47, topics/LDASamplerWithPhi.java:5-12, topics/AbortableSampler.java:3-6,
    plus getTypeTopicCounts, UncollapsedParallelLDA.java:226-234, which tests and batch builders call)."""
    methods = """setConfiguration getConfiguration addInstances addTestInstances sample setRandomSeed getNoTopics
        getNumTopics getNoTypes getCurrentIteration getZIndicators getZbar getThetaEstimate setZIndicators getDataset
        getData getDeltaStatistics getTopTypeFrequencyIndices getTypeFrequencies getCorpusSize getAlphabet getStartSeed
        getTypeMassCumSum getDocumentTopicMatrix getTypeTopicMatrix getTopicTotals getBeta getAlpha preIteration
        postIteration preSample postSample postZ preZ getLogLikelihood getHeldOutLogLikelihood abort getAbort
        getPhi setPhi getPhiMeans prePhi postPhi sampleZGivenPhi getTypeTopicCounts""".split()
    missing = [m for m in methods if not callable(getattr(L.GpuLDASampler, m, None))]
    assert not missing, missing


def test_cfg_parsing(tmp_path):
    # shape of src/main/resources/configuration/plda-cats-test.cfg
    p = tmp_path / "t.cfg"
    p.write_text(textwrap.dedent("""
        configs = ggs,pcgs
        iterations = 200
        topics = 3
        alpha = 5
        beta = 7
        seed = 2019
        exec_time = 1800 # in seconds
        compute_likelihood = false
        phi_mean_burnin = 10
        save_phi_means = false

        [ggs]
        title = LDA Grouped Gibbs Sampler
        scheme = gpu_ggs

        [pcgs]
        scheme = gpu_pcgs
        topics = 7
    """))
    c = L.LDAConfiguration.from_cfg(str(p), "ggs")
    assert (c.scheme, c.topics, c.getAlpha(), c.getBeta(), c.seed, c.iterations, c.exec_time) == \
        ("gpu_ggs", 3, 5.0, 7.0, 2019, 200, 1800.0)
    assert c.save_phi_mean is False          # the shipped cfgs write save_phi_means, the code reads save_phi_mean
    c2 = L.LDAConfiguration.from_cfg(str(p), "pcgs", topics=20)     # --topics=20 (LDACommandLineParser.java:49)
    assert (c2.scheme, c2.topics) == ("gpu_pcgs", 20)
    assert L.LDAConfiguration.from_cfg(str(p), "pcgs").topics == 7  # sub-config wins over global
    assert L.LDAConfiguration(topics=25).getAlpha() == 2.0          # ALPHA_DEFAULT = 50 / K
    with pytest.raises(KeyError):
        L.LDAConfiguration.from_cfg(str(p), "nope")


def test_instance_list_and_loader(tmp_path):
    p = tmp_path / "d.txt"
    p.write_text("docno:1\tX\tWild wild CAT cat food\ndocno:2\tX\t\ndocno:3\tY\tfood 42 lion cat\n")
    il = L.load_dataset(str(p))
    assert il.size() == 3 and il.names == ["docno:1", "docno:2", "docno:3"] and il.labels == ["X", "X", "Y"]
    assert [il.alphabet.lookupObject(i) for i in range(il.alphabet.size())] == ["wild", "cat", "food", "42", "lion"]
    off, tok = il.to_csr()
    assert off.tolist() == [0, 5, 5, 9] and tok.tolist() == [0, 0, 1, 1, 2, 2, 3, 4, 1]
    il2 = L.load_dataset(str(p), stoplist=["cat"], keep_numbers=False)
    assert il2.to_csr()[1].tolist() == [0, 0, 1, 1, 2]
    back = L.InstanceList.from_csr(off, tok, 5)
    assert back.size() == 3 and back.getNumTypes() == 5 and back.docs[2].tolist() == [2, 3, 4, 1]


def test_tokenizer_classes():
    """The four tokenizer classes LDAUtils.initTokenizer chooses from (util/LDAUtils.java:532-561)."""
    text = "ab1c foo_bar x\ty 12 z-w (q) \u00e9t\u00e9"
    # SimpleTokenizerLarge: digits, TAB and symbols are skipped without ending the token (:113-118)
    assert L.tokenize(text, keep_numbers=False) == ["abc", "foo", "bar", "xy", "z", "w", "q", "\u00e9t\u00e9"]
    # NumericAlsoTokenizer: digits build tokens
    assert L.tokenize(text, keep_numbers=True) == ["ab1c", "foo", "bar", "xy", "12", "z", "w", "q", "\u00e9t\u00e9"]
    # KeepConnectorPunctuation*: "_" (category Pc) joins (SimpleTokenizerLargeTest.java:77-97 'but_i_can')
    assert L.tokenize("yes but_i_can", keep_numbers=False, keep_connectors=True) == ["yes", "but_i_can"]
    assert "but_i_can" not in L.tokenize("yes but_i_can", keep_numbers=False)
    assert L.tokenize("the cat", stop={"the"}) == ["cat"]
    # fixed token buffer: ArrayIndexOutOfBoundsException in the reference (SimpleTokenizerLargeTest.java:48-75)
    with pytest.raises(IndexError):
        L.tokenize("abcdefghijk", max_token_buffer=10)
    assert L.tokenize("abcdefghij", max_token_buffer=10) == ["abcdefghij"]


def test_tfidf_and_rare_word_pruning(tmp_path):
    """Expected values of the reference's own test, pipe/TfIdfPipeTest.java:44-137, on its two-line corpus."""
    p = tmp_path / "tfidf.txt"
    p.write_text("docno:1\tX\tthis is a sample \ndocno:2\tX\tthis is a another another example example example\n")
    al = L.Alphabet()
    tf, df, n = L.corpus_statistics(str(p), set(), True, False, al)
    word = {al.lookupObject(i): i for i in range(al.size())}
    assert n == 2
    assert {w: tf[i] for w, i in word.items()} == {"this": 2, "is": 2, "a": 2, "sample": 1, "another": 2, "example": 3}
    assert {w: df[i] for w, i in word.items()} == {"this": 2, "is": 2, "a": 2, "sample": 1, "another": 1, "example": 1}
    ranks = L.tfidf_ranking(tf, df, n, al.size())
    assert {w: ranks.index(i) for w, i in word.items()} == \
        {"this": 5, "is": 4, "a": 3, "sample": 2, "another": 1, "example": 0}
    # loadInstancesKeep: everything from rank 3 on is stopped (TfIdfPipe.java:162-172)
    kept = L.load_dataset(str(p), tfidf_vocab_size=3)
    assert [kept.alphabet.lookupObject(i) for i in range(kept.alphabet.size())] == ["sample", "another", "example"]
    assert [d.tolist() for d in kept.docs] == [[0], [1, 1, 2, 2, 2]]
    # loadInstancesPrune: words seen fewer than rare_threshold times are stopped; first-seen order of the rest
    pruned = L.load_dataset(str(p), rare_threshold=2)
    assert [pruned.alphabet.lookupObject(i) for i in range(pruned.alphabet.size())] == ["this", "is", "a", "another", "example"]
    assert sum(len(d) for d in pruned.docs) == 11
    # a caller-supplied alphabet is shared by both passes: pruned words keep their ids (util/LDAUtils.java:250-255,295-300)
    shared = L.Alphabet()
    L.load_dataset(str(p), rare_threshold=2, alphabet=shared)
    assert shared.size() == 6 and shared.lookupObject(3) == "sample"
    # the configuration keys that drive it (ParsedLDAConfiguration.java:113,298-310,393,402-410)
    cfg = tmp_path / "c.cfg"
    cfg.write_text("scheme = gpu_ggs\nrare_threshold = 2\nkeep_numbers = true\nstoplist = stop.txt\ndataset = %s\n" % p)
    (tmp_path / "stop.txt").write_text("is\n")
    c = L.LDAConfiguration.from_cfg(str(cfg))
    il = c.loadDataset()
    assert [il.alphabet.lookupObject(i) for i in range(il.alphabet.size())] == ["this", "a", "another", "example"]


def test_cats_fixture_shape(cats):
    off, tokens = cats
    assert (len(off) - 1, int(tokens.max()) + 1, len(tokens)) == (23, 303, 7788)
    assert np.diff(off).min() == 55 and np.diff(off).max() == 916


def test_token_balanced_sharding():
    off, tokens = make_corpus(500, 100, 40, seed=3, empty_every=17)
    N = len(tokens)
    for world in (1, 2, 4, 8):
        ranges = L.shard_documents_by_tokens(off, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == 500
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [int(off[b] - off[a]) for a, b in ranges]
        assert sum(sizes) == N and max(sizes) - min(sizes) <= 2 * int(np.diff(off).max())
        pieces = [L.take_shard(off, tokens, a, b) for a, b in ranges]
        assert np.array_equal(np.concatenate([p[1] for p in pieces]), tokens)
        for (o, t, d0, t0), (a, b) in zip(pieces, ranges):
            assert o[0] == 0 and o[-1] == len(t) and d0 == a and t0 == off[a]
    # degenerate: more ranks than documents
    tiny = np.array([0, 3, 5], np.int64)
    r = L.shard_documents_by_tokens(tiny, 4)
    assert r[0][0] == 0 and r[-1][1] == 2 and all(a <= b for a, b in r)


def test_synth_corpus_is_shardable_and_deterministic():
    a_off, a_tok = L.synth_corpus(600, 300, 30.0, seed=5)
    b_off, b_tok = L.synth_corpus(600, 300, 30.0, seed=5)
    assert np.array_equal(a_tok, b_tok) and np.array_equal(a_off, b_off)
    c_off, c_tok = L.synth_corpus(200, 300, 30.0, seed=5, doc_first=250)
    assert np.array_equal(c_tok, a_tok[a_off[250]:a_off[450]])
    assert a_tok.min() >= 0 and a_tok.max() < 300 and np.diff(a_off).min() >= 1
    for d in range(0, 600, 97):     # bag of words: sorted inside a document
        seg = a_tok[a_off[d]:a_off[d + 1]]
        assert np.all(np.diff(seg) >= 0)


def test_jni_stub_type_checks_against_the_c_abi():
    """java/jni/ldagpu_jni.c (the binding for JDK 8-21) must keep compiling against include/ldagpu.h: gcc
    -fsyntax-only with a minimal stand-in for <jni.h> (no JDK in the image), warnings as errors."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "tests", "jni_stub"), os.path.join(ROOT, "java", "jni", "ldagpu_jni.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_java_shim_binds_only_declared_entry_points():
    """Every C symbol the Panama shim looks up (fn("ldagpu_...")) is declared in include/ldagpu.h."""
    src = open(os.path.join(ROOT, "java", "cc", "mallet", "topics", "GpuLDASampler.java")).read()
    used = set(re.findall(r'fn\("(ldagpu_[a-z0-9_]+)"', src))
    assert used and used <= set(L.SYMBOLS), used - set(L.SYMBOLS)


def test_row_layout_permutation_is_a_bijection_with_contiguous_lane_ownership(tmp_path):
    """common.cuh tpos / ttopic: for NT = 1, 2, 4, 8 the column of topic k is a bijection on
    [0, 128*NT), ttopic inverts it, and the float4 that lane l reads from tile j holds the lane's consecutive topics
    l*L + 4j .. l*L + 4j + 3 (L = 4*NT) -- the ownership the contract's prefix tree (oracle: draw_topic_contract_lanes)
    assumes.  Compiled with nvcc and run on the host (the functions are __host__ __device__)."""
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    src = tmp_path / "perm.cu"
    src.write_text(r'''
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace ldagpu;
int main() {
    for (int nt : {1, 2, 4, 8}) {
        const int lg = nt == 1 ? 2 : nt == 2 ? 3 : nt == 4 ? 4 : 5, L = 4 * nt, Ks = 128 * nt;
        std::vector<int> seen(Ks, 0);
        for (int k = 0; k < Ks; ++k) {
            const int p = tpos_lg(lg, k);
            if (p < 0 || p >= Ks || seen[p]++) { std::printf("not a bijection nt=%d k=%d\n", nt, k); return 1; }
            if (ttopic_lg(lg, p) != k) { std::printf("inverse nt=%d k=%d\n", nt, k); return 1; }
            const int lane = k / L, m = k % L, j = m / 4, i = m % 4;
            if (p != 128 * j + 4 * lane + i) { std::printf("ownership nt=%d k=%d\n", nt, k); return 1; }
        }
    }
    for (int k = 0; k < 30000; ++k)   // lg == 2: natural order for any K (K <= 128, K > 1024, sparse schemes)
        if (tpos_lg(2, k) != k || ttopic_lg(2, k) != k) { std::printf("identity k=%d\n", k); return 1; }
    std::puts("ok");
    return 0;
}
''')
    exe = tmp_path / "perm"
    inc = os.path.join(ROOT, "ldagroupedgibbssampler_b200", "csrc")
    r = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", inc, str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout


def test_library_carries_sm_100a_code_with_bulk_async_copies():
    """The product library is native sm_100a code, not PTX for a JIT or another architecture: every embedded cubin is
    sm_100a, the dense z-step kernels fetch their Phi^T rows with the TMA engine's bulk asynchronous copy (SASS UBLKCP,
    completion on an mbarrier: SYNCS) and the Phi kernels' peer-memory exchange uses system-scope release / acquire."""
    import shutil
    import subprocess
    from ldagroupedgibbssampler_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = _lib.library_path() if hasattr(_lib, "library_path") else os.path.join(ROOT, "ldagroupedgibbssampler_b200", "libldagpu.so")
    elfs = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True, check=True).stdout
    names = re.findall(r"ELF file\s+\d+:\s+(\S+)", elfs)
    assert names and all(n.endswith(".sm_100a.cubin") for n in names), names
    ptx = subprocess.run([cuobjdump, "-lptx", so], capture_output=True, text=True).stdout
    assert "PTX file" not in ptx, "no PTX for a JIT: the kernels are compiled for sm_100a only"
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    per_fn, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = set()
        elif cur:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                per_fn[cur].add(m.group(1))
    z = {f: ops for f, ops in per_fn.items() if "z_kernelILi" in f}     # z_kernel<NT, PCGS>, not z_kernel_big
    assert len(z) == 8, sorted(z)                              # NT in {1, 2, 4, 8} x {GGS, PCGS}
    for f, ops in z.items():
        assert "UBLKCP" in ops and "SYNCS" in ops, (f, "row fetch must be a bulk async copy on an mbarrier")
    exchange = [ops for f, ops in per_fn.items() if "phi_draw_kernel" in f or "phi_normalise_kernel" in f]
    assert exchange
    all_sass = sass
    assert ".STRONG.SYS" in all_sass, "peer-memory flags: system-scope release / acquire"
