"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and, when built, of the compiled reference pieces
(oracle/_ref/libsfref.so, oracle/_ref/libsfref_em.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arm -- never by
sailfish_b200/ (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)


def build(force=False):
    """(Re)build liboracle.so and, if /root/reference is present, oracle/_ref/*.so."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libsfref_em.so"))):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class MapOpts(C.Structure):
    _fields_ = [("max_read_occs", C.c_uint32), ("max_frag_len", C.c_uint32), ("num_frag_samples", C.c_int32),
                ("lib_format_id", C.c_int32), ("strict_intersect", C.c_int32), ("allow_orphans", C.c_int32),
                ("allow_dovetail", C.c_int32), ("ignore_compat", C.c_int32), ("enforce_compat", C.c_int32),
                ("max_interval", C.c_uint32)]

    @classmethod
    def default(cls, lib_format_id, **kw):
        o = cls(200, 1000, 10000, lib_format_id, 0, 1, 0, 0, 0, 1000)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class EMOpts(C.Structure):
    _fields_ = [("use_vb", C.c_int32), ("prior_alpha", C.c_double), ("tol", C.c_double), ("min_iter", C.c_uint32),
                ("max_iter", C.c_uint32), ("fixed_iters", C.c_uint32), ("check_cutoff", C.c_double),
                ("min_alpha", C.c_double)]

    @classmethod
    def default(cls, **kw):
        o = cls(0, 0.01, 0.01, 50, 10000, 0, 1e-2, 1e-8)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class RefBias:
    """The reference's own updateEffectiveLengths (src/SailfishUtils.cpp:611-926, compiled unmodified into oracle/_ref) on a
    ReadExperiment built from arrays; mode 1 = --biasCorrect, 2 = --gcBiasCorrect."""

    def __init__(self, mode, seqs, eff_model, read_bias, observed_gc, fld_counts, num_fwd, num_rc, gc_samp=1, classes=None, num_mapped=0,
                 use_vb=False):
        R = ref_em()
        if R is None or not hasattr(R, "ref_bias_session"):
            raise RuntimeError("oracle/_ref/libsfref_em.so was built without the bias correction")
        self.R = R
        self.T = len(seqs)
        ln = np.array([len(s) for s in seqs], np.uint32)
        eff_model = np.ascontiguousarray(eff_model, dtype=np.float64)
        rb = np.ascontiguousarray(read_bias, dtype=np.uint32); og = np.ascontiguousarray(observed_gc, dtype=np.uint32)
        fc = np.ascontiguousarray(fld_counts, dtype=np.int32)
        if classes is not None:
            rp, lab, cnt = _csr(*classes)
            self.h = R.ref_bias_session(self.T, _ptr(ln, u32p), _ptr(eff_model, f64p), b"".join(seqs), int(mode), _ptr(rb, u32p), _ptr(og, u32p),
                                        _ptr(fc, i32p), len(fc), int(num_fwd), int(num_rc), int(gc_samp), len(cnt), _ptr(rp, u64p),
                                        _ptr(lab, u32p), _ptr(cnt, u64p), int(num_mapped), int(use_vb))
        else:
            self.h = R.ref_bias_session(self.T, _ptr(ln, u32p), _ptr(eff_model, f64p), b"".join(seqs), int(mode), _ptr(rb, u32p), _ptr(og, u32p),
                                        _ptr(fc, i32p), len(fc), int(num_fwd), int(num_rc), int(gc_samp), 0, None, None, None, 0, 0)
        if not self.h:
            raise RuntimeError("ref_bias_session failed")

    def update(self, alphas, eff_in):
        alphas = np.ascontiguousarray(alphas, dtype=np.float64); eff_in = np.ascontiguousarray(eff_in, dtype=np.float64)
        out = np.zeros(self.T, np.float64)
        rc = self.R.ref_bias_update(self.h, _ptr(alphas, f64p), _ptr(eff_in, f64p), _ptr(out, f64p))
        return rc, out

    def optimize(self, tol=0.01, max_iter=10000):
        """CollapsedEMOptimizer::optimize with the session's bias mode -> (rc, estCount, EffectiveLength after the run)"""
        est = np.zeros(self.T, np.float64); mass = np.zeros(self.T, np.float64); eff = np.zeros(self.T, np.float64)
        rc = self.R.ref_em_optimize(self.h, tol, max_iter, _ptr(est, f64p), _ptr(mass, f64p))
        self.R.ref_txp_eff_lens(self.h, _ptr(eff, f64p))
        return rc, est, eff

    def fld(self, n):
        cdf = np.zeros(n, np.float32)
        mx = self.R.ref_bias_fld(self.h, _ptr(cdf, f32p), n)
        return cdf, int(mx)

    def __del__(self):
        try:
            self.R.ref_em_free(self.h)
        except Exception:
            pass


F64_ROW_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, f64p, C.c_size_t)
I32_ROW_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, i32p, C.c_size_t)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_parse_libtype.argtypes = [C.c_char_p]
        L.orc_compat_single.argtypes = [C.c_int, C.c_int32, C.c_int, C.c_int]
        L.orc_compat_paired.argtypes = [C.c_int, C.c_int]
        L.orc_hit_type.argtypes = [C.c_int32, C.c_int, C.c_uint32, C.c_int32, C.c_int, C.c_uint32, C.c_int]
        L.orc_index_build.restype = C.c_void_p
        L.orc_index_build.argtypes = [C.c_char_p, u64p, u32p, C.c_uint32, C.c_int, C.c_int]
        L.orc_index_from_arrays.restype = C.c_void_p
        L.orc_index_from_arrays.argtypes = [u64p, C.c_uint64, u32p, C.c_uint32, C.c_int, u32p, C.c_uint64]
        L.orc_index_from_table.restype = C.c_void_p
        L.orc_index_from_table.argtypes = [u64p, C.c_uint64, u32p, C.c_uint32, C.c_int, u32p, u32p, C.c_uint64, u64p, C.c_uint64]
        L.orc_index_free.argtypes = [C.c_void_p]
        for f in ("orc_index_n_sa", "orc_index_n_kmers", "orc_index_text_len"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_index_export.argtypes = [C.c_void_p, u32p, u32p, u64p, u32p, u32p]
        L.orc_index_export_text.argtypes = [C.c_void_p, u64p]
        L.orc_run_create.restype = C.c_void_p
        L.orc_run_create.argtypes = [C.c_void_p, C.POINTER(MapOpts)]
        L.orc_run_free.argtypes = [C.c_void_p]
        L.orc_run_keep_labels.argtypes = [C.c_void_p, C.c_int]
        L.orc_map_batch.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_char_p, u64p, C.c_uint64, C.c_int]
        L.orc_map_finish.argtypes = [C.c_void_p, u64p, u32p, u64p, u64p]
        L.orc_run_set_bias.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int32]
        L.orc_map_finish_bias.argtypes = [C.c_void_p, u32p, u32p]
        L.orc_eq_export.argtypes = [C.c_void_p, u64p, u32p, u64p]
        L.orc_map_work.argtypes = [C.c_void_p, u64p]
        L.orc_last_label.argtypes = [C.c_void_p, C.c_uint64, u32p, C.c_int]
        L.orc_eff_lens.argtypes = [u32p, C.c_uint32, u32p, C.c_uint32, C.c_int32, C.c_int, C.c_int, C.c_double,
                                   C.c_double, f64p]
        L.orc_update_eff_lens.argtypes = [C.c_int, C.c_uint32, C.c_char_p, u64p, u32p, f64p, f64p, f64p, C.c_int64, C.c_int64, u32p, u32p,
                                          u32p, C.c_uint32, C.c_uint32, f64p]
        L.orc_fld_cdf.restype = C.c_uint32
        L.orc_fld_cdf.argtypes = [u32p, C.c_uint32, f32p, C.c_uint32, u32p]
        L.orc_em_run_bias.argtypes = [C.c_uint32, C.c_uint64, u64p, u32p, u64p, f64p, C.c_uint64, C.POINTER(EMOpts), C.c_int, C.c_int,
                                      C.c_char_p, u64p, u32p, C.c_int64, C.c_int64, u32p, u32p, u32p, C.c_uint32, C.c_uint32, f64p, f64p, u32p, f64p]
        L.orc_em_run.argtypes = [C.c_uint32, C.c_uint64, u64p, u32p, u64p, f64p, C.c_uint64, C.POINTER(EMOpts), C.c_int,
                                 f64p, u32p, f64p]
        L.orc_tpm.argtypes = [C.c_uint32, f64p, f64p, C.c_uint64, f64p]
        L.orc_digamma.restype = C.c_double
        L.orc_digamma.argtypes = [C.c_double]
        L.orc_bootstrap.argtypes = [C.c_uint32, C.c_uint64, u64p, u32p, u64p, f64p, C.POINTER(EMOpts), C.c_uint32,
                                    C.c_uint64, F64_ROW_CB, C.c_void_p]
        L.orc_bootstrap_em.argtypes = [C.c_uint32, C.c_uint64, u64p, u32p, u64p, f64p, C.POINTER(EMOpts), f64p, u32p]
        L.orc_gibbs.argtypes = [C.c_uint32, C.c_uint64, u64p, u32p, u64p, f64p, f64p, C.c_uint64, C.c_uint32, C.c_uint64,
                                I32_ROW_CB, C.c_void_p]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------------------------
def xxh64(data, seed=0):
    b = bytes(data)
    return lib().orc_xxh64(b, len(b), seed)


def parse_libtype(s):
    return lib().orc_parse_libtype(s.encode())


def fmt_id(rtype, orient, strand):
    return (rtype & 1) | ((orient & 3) << 1) | ((strand & 7) << 3)


class Index:
    """Oracle index over a list of transcript sequences (mapping spec v1)."""

    def __init__(self, seqs=None, k=31, handle=None, txp_len=None):
        L = lib()
        if handle is not None:
            self.h, self.txp_len, self.k = handle, txp_len, k
            return
        self.k = k
        seq = b"".join(s if isinstance(s, bytes) else s.encode() for s in seqs)
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        offs = np.zeros(len(seqs), dtype=np.uint64)
        if len(seqs) > 1:
            offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
        self.txp_len = lens
        self.h = L.orc_index_build(seq, _ptr(offs, u64p), _ptr(lens, u32p), len(seqs), k, 1)
        if not self.h:
            raise RuntimeError("orc_index_build failed")

    @classmethod
    def from_text(cls, seq, txp_off, txp_len, k=31, n_threads=1):
        """index over transcripts given as one uint8 text + offsets + lengths (no per-transcript Python objects), built on n_threads"""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offs = np.ascontiguousarray(txp_off, dtype=np.uint64); lens = np.ascontiguousarray(txp_len, dtype=np.uint32)
        h = lib().orc_index_build(seq.ctypes.data_as(C.c_char_p), _ptr(offs, u64p), _ptr(lens, u32p), len(lens), k, int(n_threads))
        if not h:
            raise RuntimeError("orc_index_build failed")
        return cls(k=k, handle=h, txp_len=lens)

    @classmethod
    def from_arrays(cls, words, text_len, txp_len, k, sa_pos):
        L = lib()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        txp_len = np.ascontiguousarray(txp_len, dtype=np.uint32)
        sa_pos = np.ascontiguousarray(sa_pos, dtype=np.uint32)
        h = L.orc_index_from_arrays(_ptr(words, u64p), int(text_len), _ptr(txp_len, u32p), len(txp_len), k,
                                    _ptr(sa_pos, u32p), len(sa_pos))
        return cls(k=k, handle=h, txp_len=txp_len)

    @classmethod
    def from_table(cls, words, text_len, txp_len, k, sa_pos, sa_tid, table):
        """index built elsewhere (the GPU): packed text, suffix order with transcript ids and the k-mer table"""
        L = lib()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        txp_len = np.ascontiguousarray(txp_len, dtype=np.uint32)
        sa_pos = np.ascontiguousarray(sa_pos, dtype=np.uint32); sa_tid = np.ascontiguousarray(sa_tid, dtype=np.uint32)
        table = np.ascontiguousarray(table, dtype=np.uint64)
        h = L.orc_index_from_table(_ptr(words, u64p), int(text_len), _ptr(txp_len, u32p), len(txp_len), k, _ptr(sa_pos, u32p),
                                   _ptr(sa_tid, u32p), len(sa_pos), _ptr(table, u64p), table.size // 2)
        return cls(k=k, handle=h, txp_len=txp_len)

    def export(self):
        L = lib()
        n, m = L.orc_index_n_sa(self.h), L.orc_index_n_kmers(self.h)
        sa_pos = np.empty(n, np.uint32); sa_tid = np.empty(n, np.uint32)
        kmers = np.empty(m, np.uint64); lb = np.empty(m, np.uint32); cnt = np.empty(m, np.uint32)
        L.orc_index_export(self.h, _ptr(sa_pos, u32p), _ptr(sa_tid, u32p), _ptr(kmers, u64p), _ptr(lb, u32p), _ptr(cnt, u32p))
        return dict(sa_pos=sa_pos, sa_tid=sa_tid, kmers=kmers, lb=lb, cnt=cnt)

    def text_words(self):
        L = lib()
        n = L.orc_index_text_len(self.h)
        w = np.zeros(n // 32 + 2, np.uint64)
        L.orc_index_export_text(self.h, _ptr(w, u64p))
        return w, n

    def __del__(self):
        try:
            if self.h:
                lib().orc_index_free(self.h)
        except Exception:
            pass


def pack_reads(reads):
    """list of str/bytes -> (bytes blob, uint64 offsets[n+1])"""
    bs = [r if isinstance(r, bytes) else r.encode() for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    return b"".join(bs), off


class Run:
    """One quantification run of the oracle mapper: map batches, then finish -> classes, counters, FLD."""

    def __init__(self, index, opts):
        self.index, self.opts = index, opts
        self.h = lib().orc_run_create(index.h, C.byref(opts))

    def set_bias(self, seq_bias=True, gc_bias=False, num_bias_samples=1000000):
        """--biasCorrect / --gcBiasCorrect sample collection while mapping (call before map_batch)"""
        lib().orc_run_set_bias(self.h, int(seq_bias), int(gc_bias), int(num_bias_samples))

    def finish_bias(self):
        """-> (read_bias[4096], observed_gc[101]), pseudo-counts of 1 included"""
        rb = np.zeros(4096, np.uint32); og = np.zeros(101, np.uint32)
        rc = lib().orc_map_finish_bias(self.h, _ptr(rb, u32p), _ptr(og, u32p))
        if rc:
            raise RuntimeError("set_bias was not called")
        return rb, og

    def keep_labels(self, on=True):
        lib().orc_run_keep_labels(self.h, int(on))

    def map_batch(self, bases1, off1, bases2=None, off2=None, n_threads=1):
        n = len(off1) - 1
        off1 = np.ascontiguousarray(off1, dtype=np.uint64)
        b1 = bases1.tobytes() if isinstance(bases1, np.ndarray) else bases1
        b2 = None
        if bases2 is not None:
            off2 = np.ascontiguousarray(off2, dtype=np.uint64)
            b2 = bases2.tobytes() if isinstance(bases2, np.ndarray) else bases2
        rc = lib().orc_map_batch(self.h, b1, _ptr(off1, u64p), b2, _ptr(off2, u64p) if b2 is not None else None, n, n_threads)
        if rc != 0:
            raise RuntimeError("orc_map_batch rc=%d" % rc)

    def finish(self):
        L = lib()
        counters = np.zeros(6, np.uint64)
        fld = np.zeros(self.opts.max_frag_len, np.uint32)
        E = C.c_uint64(); nnz = C.c_uint64()
        L.orc_map_finish(self.h, _ptr(counters, u64p), _ptr(fld, u32p), C.byref(E), C.byref(nnz))
        row_ptr = np.zeros(E.value + 1, np.uint64); labels = np.zeros(max(nnz.value, 1), np.uint32)
        counts = np.zeros(max(E.value, 1), np.uint64)
        L.orc_eq_export(self.h, _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p))
        return dict(counters=counters, fld=fld, row_ptr=row_ptr, labels=labels[:nnz.value], counts=counts[:E.value])

    def work(self):
        w = np.zeros(3, np.uint64)
        lib().orc_map_work(self.h, _ptr(w, u64p))
        return w

    def last_label(self, i):
        buf = np.zeros(256, np.uint32)
        n = lib().orc_last_label(self.h, i, _ptr(buf, u32p), 256)
        return None if n < 0 else buf[:n].copy()

    def __del__(self):
        try:
            lib().orc_run_free(self.h)
        except Exception:
            pass


def eff_lens(txp_len, fld_hist, max_frag_len=1000, num_frag_samples=10000, single_end=False, mode=0,
             prior_mean=200.0, prior_sd=80.0):
    txp_len = np.ascontiguousarray(txp_len, dtype=np.uint32)
    out = np.zeros(len(txp_len), np.float64)
    fh = np.ascontiguousarray(fld_hist, dtype=np.uint32) if fld_hist is not None else None
    lib().orc_eff_lens(_ptr(txp_len, u32p), len(txp_len), _ptr(fh, u32p), max_frag_len, num_frag_samples,
                       int(single_end), mode, prior_mean, prior_sd, _ptr(out, f64p))
    return out


def update_eff_lens(mode, seqs, eff_model, eff_in, alphas, num_fwd, num_rc, read_bias, observed_gc, fld_counts, gc_samp=1):
    """sailfish::utils::updateEffectiveLengths restated (oracle/orc_bias.cpp).  mode 1 = --biasCorrect, 2 = --gcBiasCorrect;
    seqs: list of bytes (A/C/G/T).  -> (rc, eff_out)"""
    T = len(seqs)
    ln = np.array([len(s) for s in seqs], np.uint32)
    off = np.zeros(T, np.uint64); off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
    cat = b"".join(seqs)
    eff_model = np.ascontiguousarray(eff_model, dtype=np.float64); eff_in = np.ascontiguousarray(eff_in, dtype=np.float64)
    alphas = np.ascontiguousarray(alphas, dtype=np.float64)
    rb = np.ascontiguousarray(read_bias, dtype=np.uint32); og = np.ascontiguousarray(observed_gc, dtype=np.uint32)
    fc = np.ascontiguousarray(fld_counts, dtype=np.uint32)
    assert len(rb) == 4096 and len(og) == 101
    out = np.zeros(T, np.float64)
    rc = lib().orc_update_eff_lens(int(mode), T, cat, _ptr(off, u64p), _ptr(ln, u32p), _ptr(eff_model, f64p), _ptr(eff_in, f64p),
                                   _ptr(alphas, f64p), int(num_fwd), int(num_rc), _ptr(rb, u32p), _ptr(og, u32p), _ptr(fc, u32p),
                                   len(fc), int(gc_samp), _ptr(out, f64p))
    return rc, out


def fld_cdf(fld_counts):
    """EmpiricalDistribution over fragment-length counts -> (float32 cdf table, maxValue)"""
    fc = np.ascontiguousarray(fld_counts, dtype=np.uint32)
    cdf = np.zeros(len(fc) + 1, np.float32)
    mx = C.c_uint32()
    n = lib().orc_fld_cdf(_ptr(fc, u32p), len(fc), _ptr(cdf, f32p), len(cdf), C.byref(mx))
    return cdf[:n].copy(), int(mx.value)


def em_run_bias(mode, seqs, row_ptr, labels, counts, eff, num_mapped, num_fwd, num_rc, read_bias, observed_gc, fld_counts, gc_samp=1,
                opts=None, n_threads=1):
    """optimize() with --biasCorrect (mode 1) / --gcBiasCorrect (mode 2): effective lengths recomputed at iterations 50 / 500 / 1000.
    -> (rc, alphas, eff_out, iters, max_rel_diff)"""
    opts = opts or EMOpts.default()
    T = len(seqs)
    row_ptr, labels, counts = _csr(row_ptr, labels, counts)
    ln = np.array([len(s) for s in seqs], np.uint32)
    off = np.zeros(T, np.uint64); off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
    eff = np.ascontiguousarray(eff, dtype=np.float64)
    rb = np.ascontiguousarray(read_bias, dtype=np.uint32); og = np.ascontiguousarray(observed_gc, dtype=np.uint32)
    fc = np.ascontiguousarray(fld_counts, dtype=np.uint32)
    alphas = np.zeros(T, np.float64); eff_out = np.zeros(T, np.float64)
    iters = C.c_uint32(); mrd = C.c_double()
    rc = lib().orc_em_run_bias(T, len(counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p), _ptr(eff, f64p), int(num_mapped),
                               C.byref(opts), n_threads, int(mode), b"".join(seqs), _ptr(off, u64p), _ptr(ln, u32p), int(num_fwd), int(num_rc),
                               _ptr(rb, u32p), _ptr(og, u32p), _ptr(fc, u32p), len(fc), int(gc_samp), _ptr(alphas, f64p), _ptr(eff_out, f64p),
                               C.byref(iters), C.byref(mrd))
    return rc, alphas, eff_out, iters.value, mrd.value


def _csr(row_ptr, labels, counts):
    return (np.ascontiguousarray(row_ptr, dtype=np.uint64), np.ascontiguousarray(labels, dtype=np.uint32),
            np.ascontiguousarray(counts, dtype=np.uint64))


def em_run(n_txp, row_ptr, labels, counts, eff, num_mapped, opts=None, n_threads=1):
    opts = opts or EMOpts.default()
    row_ptr, labels, counts = _csr(row_ptr, labels, counts)
    eff = np.ascontiguousarray(eff, dtype=np.float64)
    alphas = np.zeros(n_txp, np.float64)
    iters = C.c_uint32(); mrd = C.c_double()
    rc = lib().orc_em_run(n_txp, len(counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p), _ptr(eff, f64p),
                          int(num_mapped), C.byref(opts), n_threads, _ptr(alphas, f64p), C.byref(iters), C.byref(mrd))
    return rc, alphas, iters.value, mrd.value


def bootstrap_em(n_txp, row_ptr, labels, samp_counts, eff, opts=None):
    opts = opts or EMOpts.default()
    row_ptr, labels, samp_counts = _csr(row_ptr, labels, samp_counts)
    eff = np.ascontiguousarray(eff, dtype=np.float64)
    alphas = np.zeros(n_txp, np.float64)
    iters = C.c_uint32()
    rc = lib().orc_bootstrap_em(n_txp, len(samp_counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(samp_counts, u64p),
                                _ptr(eff, f64p), C.byref(opts), _ptr(alphas, f64p), C.byref(iters))
    return rc, alphas, iters.value


def tpm(alphas, eff, num_mapped):
    alphas = np.ascontiguousarray(alphas, dtype=np.float64); eff = np.ascontiguousarray(eff, dtype=np.float64)
    out = np.zeros(len(alphas), np.float64)
    lib().orc_tpm(len(alphas), _ptr(alphas, f64p), _ptr(eff, f64p), int(num_mapped), _ptr(out, f64p))
    return out


def bootstrap(n_txp, row_ptr, labels, counts, eff, n_boot, seed=1, opts=None):
    opts = opts or EMOpts.default()
    row_ptr, labels, counts = _csr(row_ptr, labels, counts)
    eff = np.ascontiguousarray(eff, dtype=np.float64)
    rows = []
    cb = F64_ROW_CB(lambda u, p, n: (rows.append(np.ctypeslib.as_array(p, shape=(n,)).copy()), 0)[1])
    rc = lib().orc_bootstrap(n_txp, len(counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p), _ptr(eff, f64p),
                             C.byref(opts), n_boot, seed, cb, None)
    return rc, np.array(rows)


def gibbs(n_txp, row_ptr, labels, counts, eff, masses, num_mapped, n_samples, seed=1):
    row_ptr, labels, counts = _csr(row_ptr, labels, counts)
    eff = np.ascontiguousarray(eff, dtype=np.float64); masses = np.ascontiguousarray(masses, dtype=np.float64)
    rows = []
    cb = I32_ROW_CB(lambda u, p, n: (rows.append(np.ctypeslib.as_array(p, shape=(n,)).copy()), 0)[1])
    rc = lib().orc_gibbs(n_txp, len(counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p), _ptr(eff, f64p),
                         _ptr(masses, f64p), int(num_mapped), n_samples, seed, cb, None)
    return rc, np.array(rows)


# ---------------------------------------------------------------------------------------------------------------
# compiled reference pieces (oracle/_ref) -- present in this container and shipped prebuilt to the GPU box
_ref = None
_ref_em = None


def ref():
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libsfref.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_xxh64.restype = C.c_uint64
        R.ref_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        R.ref_tgroup_hash.restype = C.c_uint64
        R.ref_tgroup_hash.argtypes = [u32p, C.c_uint32]
        R.ref_eqb_create.restype = C.c_void_p
        R.ref_eqb_free.argtypes = [C.c_void_p]
        R.ref_eqb_add.argtypes = [C.c_void_p, u32p, C.c_uint32]
        R.ref_eqb_finish.restype = C.c_uint64
        R.ref_eqb_finish.argtypes = [C.c_void_p, u64p]
        R.ref_eqb_export.argtypes = [C.c_void_p, u64p, u32p, u64p, f64p]
        R.ref_empdist.restype = C.c_float
        R.ref_empdist.argtypes = [u32p, u32p, C.c_uint32, f32p, C.c_uint32, u32p, u32p]
        _ref = R
    return _ref


def ref_em():
    global _ref_em
    if _ref_em is None:
        p = os.path.join(_HERE, "_ref", "libsfref_em.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_em_session.restype = C.c_void_p
        R.ref_em_session.argtypes = [C.c_uint32, u32p, f64p, C.c_uint64, u64p, u32p, u64p, C.c_uint64, C.c_int, C.c_int]
        R.ref_em_free.argtypes = [C.c_void_p]
        R.ref_em_eq_order.restype = C.c_uint64
        R.ref_em_eq_order.argtypes = [C.c_void_p, u64p, u32p, u64p]
        R.ref_em_optimize.argtypes = [C.c_void_p, C.c_double, C.c_uint32, f64p, f64p]
        R.ref_em_bootstraps.argtypes = [C.c_void_p, C.c_double, C.c_uint32, f64p]
        R.ref_em_gibbs.argtypes = [C.c_void_p, C.c_uint32, i32p]
        if hasattr(R, "ref_bias_session"):
            R.ref_bias_session.restype = C.c_void_p
            R.ref_bias_session.argtypes = [C.c_uint32, u32p, f64p, C.c_char_p, C.c_int, u32p, u32p, i32p, C.c_uint32, C.c_int64, C.c_int64,
                                           C.c_uint32, C.c_uint64, u64p, u32p, u64p, C.c_uint64, C.c_int]
            R.ref_bias_update.argtypes = [C.c_void_p, f64p, f64p, f64p]
            R.ref_bias_fld.restype = C.c_uint32
            R.ref_bias_fld.argtypes = [C.c_void_p, f32p, C.c_uint32]
            R.ref_txp_eff_lens.argtypes = [C.c_void_p, f64p]
            R.ref_readbias_update.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
            R.ref_gc_frac.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        _ref_em = R
    return _ref_em


class RefEM:
    """The reference's own CollapsedEMOptimizer / CollapsedGibbsSampler on a ReadExperiment built from arrays."""

    def __init__(self, txp_len, eff, row_ptr, labels, counts, num_mapped, use_vb=False, n_boot=0):
        R = ref_em()
        if R is None:
            raise RuntimeError("oracle/_ref/libsfref_em.so is not built")
        self.R = R
        self.T = len(txp_len)
        txp_len = np.ascontiguousarray(txp_len, dtype=np.uint32); eff = np.ascontiguousarray(eff, dtype=np.float64)
        row_ptr, labels, counts = _csr(row_ptr, labels, counts)
        self.E, self.nnz, self.n_boot = len(counts), len(labels), n_boot
        self.h = R.ref_em_session(self.T, _ptr(txp_len, u32p), _ptr(eff, f64p), self.E, _ptr(row_ptr, u64p), _ptr(labels, u32p),
                                  _ptr(counts, u64p), int(num_mapped), int(use_vb), int(n_boot))

    def eq_order(self):
        rp = np.zeros(self.E + 1, np.uint64); lab = np.zeros(max(self.nnz, 1), np.uint32); cnt = np.zeros(max(self.E, 1), np.uint64)
        self.R.ref_em_eq_order(self.h, _ptr(rp, u64p), _ptr(lab, u32p), _ptr(cnt, u64p))
        return rp, lab[:self.nnz], cnt[:self.E]

    def optimize(self, tol=0.01, max_iter=10000):
        est = np.zeros(self.T, np.float64); mass = np.zeros(self.T, np.float64)
        rc = self.R.ref_em_optimize(self.h, tol, max_iter, _ptr(est, f64p), _ptr(mass, f64p))
        return rc, est, mass

    def bootstraps(self, tol=0.01, max_iter=10000):
        out = np.zeros((self.n_boot, self.T), np.float64)
        rc = self.R.ref_em_bootstraps(self.h, tol, max_iter, _ptr(out, f64p))
        return rc, out

    def gibbs(self, n_samples):
        out = np.zeros((n_samples, self.T), np.int32)
        rc = self.R.ref_em_gibbs(self.h, n_samples, _ptr(out, i32p))
        return rc, out

    def __del__(self):
        try:
            self.R.ref_em_free(self.h)
        except Exception:
            pass
