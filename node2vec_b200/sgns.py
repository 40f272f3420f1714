"""Host side of the SGNS half (K3/K4): a gensim-3.8-Word2Vec-shaped model whose training
runs in the CUDA library.

``Word2Vec`` mirrors the slice of gensim 3.8's ``Word2Vec`` / ``KeyedVectors`` API that the
reference touches (embedding.py:126, 135-136, 151, 157, 163, 170, 177): constructor keyword
names (``size, window, min_count, alpha, iter, seed, batch_words, negative, workers, sg,
sample, min_alpha, ns_exponent, hs``), ``model.wv.vocab`` (insertion order = first
appearance in the corpus), ``model.wv[token]``, ``save`` / ``load`` and word2vec text
format.  Tokens are the decimal strings of vertex ids, exactly what
``np.array(walks).astype(str)`` feeds gensim (embedding.py:125).

Only the configuration on the reference's hot path is trained: ``sg=1, hs=0, negative>0``
(skip-gram with negative sampling).  ``negative=0`` with ``hs=0`` performs no updates, as in
gensim 3.8 (the reference's own defaults, constants.py:50-68).  CBOW / hierarchical softmax
raise NotImplementedError.
"""
import ctypes as C
import pickle
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib, dist as n2v_dist

_INT64_MAX = np.iinfo(np.int64).max
NEG_CHUNK = 8192          # include/n2v_b200.h: N2V_NEG_CHUNK


def neg_top_entries(n_rows: int) -> int:
    """n2v_neg_top_entries(): entries of the top-level negative table (0 = single level)."""
    return (n_rows + NEG_CHUNK - 1) // NEG_CHUNK if n_rows > 65536 else 0


class Vocab(object):
    """gensim.models.keyedvectors.Vocab look-alike."""

    __slots__ = ("count", "index", "sample_int")

    def __init__(self, count=0, index=0, sample_int=0):
        self.count, self.index, self.sample_int = count, index, sample_int

    def __repr__(self):
        return f"Vocab(count:{self.count}, index:{self.index}, sample_int:{self.sample_int})"


class KeyedVectors(object):
    """The ``model.wv`` object: token -> vector."""

    def __init__(self, vector_size: int):
        self.vector_size = vector_size
        self._vocab: Optional[Dict[str, Vocab]] = {}
        self._index2word: Optional[list] = []
        self._lazy = None       # (ids in first-appearance order, rank by count, counts, keep thresholds)
        self._vectors = np.zeros((0, vector_size), dtype=np.float32)
        self._vectors_fn = None  # device -> host copy of the table, run on first access after training

    @property
    def vectors(self) -> np.ndarray:
        if self._vectors_fn is not None:
            self._vectors, self._vectors_fn = self._vectors_fn(), None
        return self._vectors

    @vectors.setter
    def vectors(self, value):
        self._vectors, self._vectors_fn = value, None

    # The dict of Vocab objects costs ~1 us per vertex to build in Python; it is materialised on
    # first access only (training never needs it).
    def _materialise(self):
        if self._lazy is not None:
            ids, rank, counts, keep = (t.cpu().numpy() for t in self._lazy)
            keep = keep.view(np.uint32)
            self._lazy = None
            self._vocab = {str(int(v)): Vocab(int(c), int(r), 2 ** 32 if k == 0xFFFFFFFF else int(k))
                           for v, r, c, k in zip(ids.tolist(), rank.tolist(), counts.tolist(), keep.tolist())}
            order = ids[np.argsort(rank)]
            self._index2word = [str(int(v)) for v in order.tolist()]

    @property
    def vocab(self) -> Dict[str, Vocab]:
        self._materialise()
        return self._vocab

    @vocab.setter
    def vocab(self, value):
        self._lazy, self._vocab = None, value

    @property
    def index2word(self):
        self._materialise()
        return self._index2word

    @index2word.setter
    def index2word(self, value):
        self._index2word = value

    def __getitem__(self, token):
        if isinstance(token, (list, tuple)):
            return np.vstack([self[t] for t in token])
        return self.vectors[self.vocab[str(token)].index]

    get_vector = word_vec = __getitem__

    def __contains__(self, token):
        return str(token) in self.vocab

    def __len__(self):
        return len(self.index2word)

    def save_word2vec_format(self, fname: str) -> None:
        """word2vec text format, most frequent token first (gensim's order)."""
        with open(fname, "w") as f:
            f.write(f"{len(self.index2word)} {self.vector_size}\n")
            for word in self.index2word:
                row = self.vectors[self.vocab[word].index]
                f.write(word + " " + " ".join(str(x) for x in row) + "\n")

    @classmethod
    def load_word2vec_format(cls, fname: str) -> "KeyedVectors":
        with open(fname) as f:
            n, d = (int(x) for x in f.readline().split())
            kv = cls(d)
            kv.vectors = np.zeros((n, d), dtype=np.float32)
            for i in range(n):
                parts = f.readline().rstrip("\n").split(" ")
                kv.vocab[parts[0]] = Vocab(count=n - i, index=i)
                kv.index2word.append(parts[0])
                kv.vectors[i] = np.asarray(parts[1:], dtype=np.float32)
        return kv


def _walk_matrix(sentences, device) -> torch.Tensor:
    """Rectangular int32 token matrix on the device from: a torch tensor, a numpy matrix, or
    lists of int / decimal-string tokens (what the reference hands gensim)."""
    if isinstance(sentences, torch.Tensor):
        return sentences.to(device=device, dtype=torch.int32)
    arr = np.asarray(sentences)
    if arr.dtype.kind in "US" or arr.dtype == object:
        arr = arr.astype(np.int64)
    if arr.ndim != 2:
        raise ValueError("walks must be rectangular (every walk the same length)")
    return torch.as_tensor(np.ascontiguousarray(arr.astype(np.int32, copy=False)), device=device)


class Word2Vec(object):
    def __init__(self, sentences=None, size=100, alpha=0.025, window=5, min_count=5, sample=1e-3, seed=1,
                 workers=3, min_alpha=0.0001, sg=0, hs=0, negative=5, ns_exponent=0.75, iter=5,
                 batch_words=10000, atomic_updates=True, process_group=None, sync_every=1,
                 share_negatives=False, **ignored):
        _lib.require_cuda()
        if hs:
            raise NotImplementedError("hierarchical softmax (hs=1) is outside the SGNS hot path")
        if negative and not sg:
            raise NotImplementedError("CBOW (sg=0) with negative sampling is outside the SGNS hot path; pass sg=1")
        if size % 4 != 0 or size < 4 or size > 1024:
            raise ValueError(f"vector size {size} must be a multiple of 4 in [4, 1024]")
        self.vector_size, self.alpha, self.window, self.min_count = int(size), float(alpha), int(window), int(min_count)
        self.sample, self.seed, self.workers, self.min_alpha = float(sample), int(seed or 0), workers, float(min_alpha)
        self.sg, self.hs, self.negative, self.ns_exponent = int(sg), int(hs), int(negative), float(ns_exponent)
        self.epochs = self.iter = int(iter)
        self.batch_words = int(batch_words)
        self.atomic_updates = bool(atomic_updates)
        # EXPERIMENTAL, off by default: draw the K negatives once per centre instead of once per pair
        # (csrc/sgns_shared.cu); a different sampling scheme with its own AUC gate, not the parity path
        self.share_negatives = bool(share_negatives)
        if self.share_negatives and (size > 128 or negative != 5 or not atomic_updates):
            raise ValueError("share_negatives needs size <= 128, negative == 5 and atomic_updates=True")
        self.process_group, self.sync_every = process_group, max(1, int(sync_every))
        self.wv = KeyedVectors(self.vector_size)
        self.corpus_count = 0
        self.train_stats: Dict[str, int] = {}
        self.syn0 = self.syn1neg = None      # device tables [n_rows, size]
        self._ids = None                     # row -> vertex id when the id space was densified
        if sentences is not None:
            # one host->device copy of the corpus serves both passes
            sentences = _walk_matrix(sentences, torch.device("cuda", torch.cuda.current_device()))
            self.build_vocab(sentences)
            self.train(sentences)

    # ------------------------------------------------------------------ vocabulary (K4)
    def build_vocab(self, sentences) -> None:
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        walks = _walk_matrix(sentences, dev)
        self.corpus_count = int(walks.shape[0])
        if walks.numel() == 0:
            raise RuntimeError("you must first build vocabulary before training the model")
        lo, hi = int(walks.min()), int(walks.max())
        if hi < 0:
            raise ValueError("no valid token in the corpus")
        self._n_rows = hi + 1
        stream = _lib.current_stream_ptr()
        counts = torch.zeros(self._n_rows, dtype=torch.int64, device=dev)
        first = torch.full((self._n_rows,), _INT64_MAX, dtype=torch.int64, device=dev)
        row_off, self._total_walks, n_rows = n2v_dist.shard_layout(self.corpus_count, self._n_rows,
                                                                   self.process_group) \
            if self.process_group is not None else (0, self.corpus_count, self._n_rows)
        if n_rows != self._n_rows:
            self._n_rows = n_rows
            counts = torch.zeros(n_rows, dtype=torch.int64, device=dev)
            first = torch.full((n_rows,), _INT64_MAX, dtype=torch.int64, device=dev)
        self._walk_offset = row_off
        _lib.check(lib.n2v_vocab_count(_lib.ptr(walks), walks.shape[0], walks.shape[1], walks.stride(0),
                                       self._n_rows, row_off * walks.shape[1], _lib.ptr(counts), _lib.ptr(first),
                                       stream), "n2v_vocab_count")
        if self.process_group is not None:
            n2v_dist.reduce_vocab(counts, first, self.process_group)
        self._keep = torch.empty(self._n_rows, dtype=torch.int32, device=dev)
        n_top = neg_top_entries(self._n_rows)
        self._neg = torch.empty((self._n_rows + n_top, 2), dtype=torch.int32, device=dev)
        scratch = torch.empty((self._n_rows + n_top) * 20 + 64, dtype=torch.uint8, device=dev)
        totals = (C.c_int64 * 2)()
        _lib.check(lib.n2v_sgns_prepare(_lib.ptr(counts), self._n_rows, self.min_count, self.sample,
                                        self.ns_exponent, _lib.ptr(self._keep), _lib.ptr(self._neg),
                                        _lib.ptr(scratch), totals, stream), "n2v_sgns_prepare")
        self._retain_total, n_vocab = int(totals[0]), int(totals[1])
        # wv.vocab: insertion order = first appearance in the corpus; .index = rank by count (desc, ties by
        # first appearance) -- computed with device sorts, kept on the device until somebody asks
        ids = torch.nonzero((counts > 0) & (counts >= self.min_count)).view(-1)
        ids = ids[torch.sort(first[ids], stable=True).indices]                 # first-appearance order
        by_count = torch.sort(-counts[ids], stable=True).indices               # stable: ties keep that order
        rank = torch.empty_like(by_count)
        rank[by_count] = torch.arange(by_count.numel(), device=dev)
        self.wv._lazy = (ids, rank, counts[ids], self._keep[ids])
        self._row_of_index = ids[by_count]
        assert n_vocab == int(ids.numel())
        # weights: reset_weights()
        self.syn0 = torch.empty((self._n_rows, self.vector_size), dtype=torch.float32, device=dev)
        self.syn1neg = torch.zeros((self._n_rows, self.vector_size), dtype=torch.float32, device=dev)
        _lib.check(lib.n2v_sgns_init(_lib.ptr(self.syn0), self._n_rows, self.vector_size,
                                     C.c_uint64(self.seed & 0xFFFFFFFFFFFFFFFF), stream), "n2v_sgns_init")
        table = (C.c_float * 1000)()
        _lib.check(lib.n2v_sgns_exp_table(table))
        self._exp = torch.as_tensor(np.frombuffer(table, dtype=np.float32).copy(), device=dev)
        self._sync_vectors()

    # ------------------------------------------------------------------ training (K3)
    def train(self, sentences, epochs: Optional[int] = None, trace_cap: int = 0,
              epoch_range: Optional[Tuple[int, int]] = None, walk_range: Optional[Tuple[int, int]] = None,
              sync: Optional[bool] = None, **ignored):
        """Run ``epochs`` (default ``iter``) epochs.  Returns (pairs trained, tokens kept).
        ``epoch_range=(a, b)`` runs only epochs a..b-1 of that ``epochs``-long schedule (learning-rate
        decay and random streams are functions of the epoch index), so training can be checkpointed
        with ``save`` between epochs and resumed after ``load`` with the remaining range.
        ``walk_range=(lo, hi)`` trains on rows lo..hi-1 only (a slice of the epoch: learning rate and
        random streams are functions of the row's global index, so slices compose to the epoch).
        ``sync`` overrides the data-parallel averaging decision for this call (None = every
        ``sync_every`` epochs and after the last one)."""
        lib = _lib.load()
        if self.syn0 is None:
            raise RuntimeError("you must first build vocabulary before training the model")
        dev = self.syn0.device
        walks = _walk_matrix(sentences, dev)
        epochs = self.epochs if epochs is None else int(epochs)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        trace = trace_alpha = None
        if trace_cap:
            trace = torch.full((trace_cap, 2 + max(self.negative, 1)), -2, dtype=torch.int32, device=dev)
            trace_alpha = torch.zeros(trace_cap, dtype=torch.float32, device=dev)
        if self.negative > 0:
            P = _lib.SgnsParams(dim=self.vector_size, window=self.window, negative=self.negative, epochs=epochs,
                                epoch=0, batch_words=self.batch_words, atomic_updates=int(self.atomic_updates),
                                alpha=self.alpha, min_alpha=self.min_alpha, seed=self.seed & 0xFFFFFFFFFFFFFFFF,
                                walk_offset=self._walk_offset, total_walks=self._total_walks)
            first, last = (0, epochs) if epoch_range is None else (int(epoch_range[0]), int(epoch_range[1]))
            if not 0 <= first <= last <= epochs:
                raise ValueError(f"epoch_range {epoch_range} must lie inside [0, {epochs}]")
            rows = walks
            if walk_range is not None:
                lo, hi = int(walk_range[0]), int(walk_range[1])
                if not 0 <= lo <= hi <= walks.shape[0]:
                    raise ValueError(f"walk_range {walk_range} must lie inside [0, {walks.shape[0]}]")
                rows = walks[lo:hi]
                P.walk_offset = self._walk_offset + lo
            for ep in range(first, last):
                P.epoch = ep
                entry = lib.n2v_sgns_train_shared if getattr(self, "share_negatives", False) else lib.n2v_sgns_train
                _lib.check(entry(_lib.ptr(rows), rows.shape[0], walks.shape[1], walks.stride(0),
                                 _lib.ptr(self._keep), _lib.ptr(self._neg), self._n_rows,
                                 _lib.ptr(self.syn0), _lib.ptr(self.syn1neg), _lib.ptr(self._exp),
                                 C.byref(P), _lib.ptr(stats), _lib.ptr(trace), _lib.ptr(trace_alpha),
                                 trace_cap, _lib.current_stream_ptr()), "n2v_sgns_train")
                due = ((ep + 1) % self.sync_every == 0 or ep + 1 == epochs) if sync is None else bool(sync)
                if self.process_group is not None and due:
                    self.average_tables()
        st = dict(zip(_lib.SGNS_STAT_NAMES, stats.cpu().tolist()))
        self.train_stats = st
        self._sync_vectors()
        if trace_cap:
            self.last_trace = (trace.cpu().numpy(), trace_alpha.cpu().numpy())
        return st["pairs"], st["tokens_kept"]

    def average_tables(self) -> None:
        """Data-parallel model averaging: NCCL sum-allreduce over NVLink, then x 1/G (n2v_scale)."""
        lib = _lib.load()

        def scale(t, f):
            _lib.check(lib.n2v_scale(_lib.ptr(t), t.numel(), f, _lib.current_stream_ptr()), "n2v_scale")
        n2v_dist.average_tables((self.syn0, self.syn1neg), self.process_group, scale)

    def _sync_vectors(self) -> None:
        # lazy: the [vocab, size] host copy is made when somebody reads wv.vectors / wv[token]
        def to_host():
            rows = self.syn0[self._row_of_index]
            nbytes = rows.numel() * rows.element_size()
            if 0 < nbytes <= (4 << 30):
                # a pinned landing buffer (torch caches it): one DMA instead of a staged copy into freshly
                # faulted pageable memory -- ~0.2 s of the 0.5 GB table read-back on BASELINE configs[2]
                host = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
                host.copy_(rows)
                return host.numpy()
            return rows.cpu().numpy()
        self.wv._vectors_fn = to_host

    # ------------------------------------------------------------------ persistence
    def save(self, fname: str) -> None:
        self.wv._materialise()
        _ = self.wv.vectors
        state = {k: v for k, v in self.__dict__.items()
                 if k not in ("syn0", "syn1neg", "_keep", "_neg", "_exp", "_row_of_index", "process_group")}
        state["syn0"] = None if self.syn0 is None else self.syn0.cpu().numpy()
        state["syn1neg"] = None if self.syn1neg is None else self.syn1neg.cpu().numpy()
        state["_row_of_index"] = None if self.syn0 is None else self._row_of_index.cpu().numpy()
        # sampling tables too, so that a loaded model can continue training (train(epoch_range=...))
        for name in ("_keep", "_neg", "_exp"):
            t = getattr(self, name, None)
            state[name] = None if t is None else t.cpu().numpy()
        with open(fname, "wb") as f:
            pickle.dump(state, f, protocol=4)

    @classmethod
    def load(cls, fname: str) -> "Word2Vec":
        with open(fname, "rb") as f:
            state = pickle.load(f)
        self = cls.__new__(cls)
        self.__dict__.update(state)
        self.process_group = None
        if state["syn0"] is not None and torch.cuda.is_available():
            dev = torch.device("cuda", torch.cuda.current_device())
            self.syn0 = torch.as_tensor(state["syn0"], device=dev)
            self.syn1neg = torch.as_tensor(state["syn1neg"], device=dev)
            self._row_of_index = torch.as_tensor(state["_row_of_index"], device=dev)
            for name in ("_keep", "_neg", "_exp"):
                if state.get(name) is not None:
                    setattr(self, name, torch.as_tensor(state[name], device=dev))
        return self
