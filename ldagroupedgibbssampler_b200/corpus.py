"""Corpus containers on the host side of the drop-in boundary.

The reference hands the sampler a MALLET ``InstanceList`` whose instances carry a ``FeatureSequence``
of type ids (reference: src/main/java/cc/mallet/util/LDAUtils.java:136-182,233-330).  This module
holds the minimum of that: an ``InstanceList`` mirror that flattens to CSR (``doc_offsets`` int64[D+1],
``tokens`` int32[N]), a reader for the reference's ``name<TAB>label<TAB>text`` files that is enough for
its bundled bag-of-words corpora, the named benchmark shapes of SURVEY section 8, and the
token-balanced document sharding of section 8(e).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np


class Alphabet:
    """First-seen-order vocabulary (MALLET Alphabet as StringList2FeatureSequence fills it)."""

    def __init__(self, entries: Optional[Iterable[str]] = None):
        self._list: List[str] = []
        self._index = {}
        for e in entries or []:
            self.lookupIndex(e)

    def lookupIndex(self, entry: str, add: bool = True) -> int:
        i = self._index.get(entry)
        if i is None:
            if not add:
                return -1
            i = len(self._list)
            self._index[entry] = i
            self._list.append(entry)
        return i

    def lookupObject(self, i: int) -> str:
        return self._list[i]

    def size(self) -> int:
        return len(self._list)

    def __len__(self) -> int:
        return len(self._list)

    def __eq__(self, other) -> bool:
        return isinstance(other, Alphabet) and self._list == other._list


@dataclass
class InstanceList:
    """Documents as arrays of type ids plus the alphabets, like MALLET's InstanceList of FeatureSequences."""

    docs: List[np.ndarray] = field(default_factory=list)
    alphabet: Alphabet = field(default_factory=Alphabet)
    names: List[str] = field(default_factory=list)
    labels: List[str] = field(default_factory=list)
    num_types: Optional[int] = None   # for synthetic corpora that have no strings

    def size(self) -> int:
        return len(self.docs)

    def getDataAlphabet(self) -> Alphabet:
        return self.alphabet

    def getNumTypes(self) -> int:
        return self.num_types if self.num_types is not None else self.alphabet.size()

    def to_csr(self) -> Tuple[np.ndarray, np.ndarray]:
        if getattr(self, "_csr", None) is not None:
            return self._csr
        lens = np.fromiter((len(d) for d in self.docs), np.int64, len(self.docs))
        off = np.zeros(len(self.docs) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        tokens = (np.concatenate([np.asarray(d, np.int32) for d in self.docs]) if off[-1] > 0
                  else np.zeros(0, np.int32))
        return off, np.ascontiguousarray(tokens, np.int32)

    @staticmethod
    def from_csr(doc_offsets: np.ndarray, tokens: np.ndarray, num_types: int) -> "InstanceList":
        il = InstanceList(num_types=int(num_types))
        il._csr = (np.ascontiguousarray(doc_offsets, np.int64), np.ascontiguousarray(tokens, np.int32))
        il.docs = _CsrDocs(*il._csr)
        return il


class _CsrDocs(Sequence):
    """Sequence view over a CSR corpus (so million-document synthetic corpora are not split into lists)."""

    def __init__(self, off, tokens):
        self.off, self.tokens = off, tokens

    def __len__(self):
        return len(self.off) - 1

    def __getitem__(self, d):
        if isinstance(d, slice):
            return [self[i] for i in range(*d.indices(len(self)))]
        return self.tokens[self.off[d]: self.off[d + 1]]


_LINE = re.compile(r"^(\S*)[\s,]*([^\t]+)[\s,]*(.*)$")   # LDAUtils.java:236
_TOKEN = re.compile(r"[^\W_]+", re.UNICODE)


def load_dataset(path: str, stoplist: Optional[Iterable[str]] = None, keep_numbers: bool = True,
                 alphabet: Optional[Alphabet] = None) -> InstanceList:
    """Read a ``name<TAB>label<TAB>text`` file (LDAUtils.java:233-330): lower-case, tokenise on
    non-alphanumerics, drop stop words, first-seen vocabulary order.  Enough for the reference's
    bundled bag-of-words corpora (cats.txt, small.txt); rare-word pruning, TF-IDF pruning and the
    connector-punctuation options of SimpleTokenizerLarge are not restated."""
    stop = set(stoplist or [])
    il = InstanceList(alphabet=alphabet or Alphabet())
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if not line.strip():
                continue
            m = _LINE.match(line)
            if not m:
                continue
            name, label, text = m.group(1), m.group(2).strip(), m.group(3)
            ids = []
            for tok in _TOKEN.findall(text.lower()):
                if tok in stop or (not keep_numbers and tok.isdigit()):
                    continue
                ids.append(il.alphabet.lookupIndex(tok))
            il.docs.append(np.asarray(ids, np.int32))
            il.names.append(name)
            il.labels.append(label)
    return il


# --------------------------------------------------------------------------------------------
# named shapes (SURVEY section 8 size table; reference: src/main/resources/datasets/README.txt:3-31)
# --------------------------------------------------------------------------------------------
SHAPES = {
    # name: (D, V, mean document length, K, scheme, alpha, beta)
    "nips": dict(D=1500, V=12419, mean_len=1267.0, K=100, scheme="gpu_ggs", alpha=1.0, beta=0.01),
    "enron": dict(D=39861, V=28102, mean_len=161.0, K=400, scheme="gpu_pcgs", alpha=50.0 / 400, beta=0.01),
    "pubmed": dict(D=8200000, V=141043, mean_len=90.0, K=1000, scheme="gpu_ggs", alpha=50.0 / 1000, beta=0.01),
}


def shard_documents_by_tokens(doc_offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous document ranges with (nearly) equal token counts, one per rank (SURVEY 8e:
    balance by tokens, not by document count as the reference's EvenSplitBatchBuilder does,
    topics/randomscan/document/EvenSplitBatchBuilder.java:30-44).  Returns [(d0, d1)] per rank."""
    off = np.asarray(doc_offsets, np.int64)
    D, N = len(off) - 1, int(off[-1])
    cuts = [0]
    for r in range(1, world):
        target = N * r // world
        d = int(np.searchsorted(off, target, side="left"))
        d = min(max(d, cuts[-1]), D)
        cuts.append(d)
    cuts.append(D)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def take_shard(doc_offsets: np.ndarray, tokens: np.ndarray, d0: int, d1: int):
    """Local CSR of documents [d0, d1) plus the global bases the library keys its counters with."""
    off = np.asarray(doc_offsets, np.int64)
    t0, t1 = int(off[d0]), int(off[d1])
    return (off[d0: d1 + 1] - t0).astype(np.int64), np.ascontiguousarray(tokens[t0:t1], np.int32), d0, t0
