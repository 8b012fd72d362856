"""ctypes binding of libsfb200.so (include/sfb200.h) -- what tests/, bench.py and __graft_entry__ call.

There is no CPU fallback: if the library is missing it is built with nvcc; if no CUDA device is usable every compute
call raises.  Nothing here imports oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFB200_LIB", os.path.join(_HERE, "libsfb200.so"))

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)

ERRORS = {-1: "ENODEV", -2: "ECUDA", -3: "EINVAL", -4: "ENOACTIVE", -5: "ESMALLSUM", -6: "EFULL", -7: "ENCCL", -8: "ECALLBACK"}


class Sfb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sfb200 error %d (%s): %s" % (code, ERRORS.get(code, "?"), msg))
        self.code = code


class MapOpts(C.Structure):
    """sfb200_map_opts: the SailfishOpts members processReadsQuasi reads (reference include/SailfishOpts.hpp:9-41)."""
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
    """sfb200_em_opts: constants hard-coded in the reference (CollapsedEMOptimizer.cpp:716,786,810-811; SailfishQuantify.cpp:1343)."""
    _fields_ = [("use_vb", C.c_int32), ("prior_alpha", C.c_double), ("tol", C.c_double), ("min_iter", C.c_uint32),
                ("max_iter", C.c_uint32), ("fixed_iters", C.c_uint32), ("check_cutoff", C.c_double),
                ("min_alpha", C.c_double)]

    @classmethod
    def default(cls, **kw):
        o = cls(0, 0.01, 0.01, 50, 10000, 0, 1e-2, 1e-8)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class BiasModel(C.Structure):
    """sfb200_bias_model"""
    _fields_ = [("mode", C.c_int32), ("gc_samp", C.c_uint32), ("num_fwd", C.c_int64), ("num_rc", C.c_int64), ("read_bias", u32p),
                ("observed_gc", u32p), ("fld_cdf", C.POINTER(C.c_float)), ("n_cdf", C.c_uint32), ("fld_max", C.c_uint32)]


F64_ROW_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, f64p, C.c_size_t)
I32_ROW_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, i32p, C.c_size_t)

# every entry point include/sfb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "sfb200_version": (C.c_int, []),
    "sfb200_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sfb200_ctx_destroy": (None, [C.c_void_p]),
    "sfb200_last_error": (C.c_char_p, [C.c_void_p]),
    "sfb200_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sfb200_ctx_sync": (C.c_int, [C.c_void_p]),
    "sfb200_launch_count": (C.c_uint64, [C.c_void_p]),
    "sfb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "sfb200_bind_host_near_device": (C.c_int, [C.c_int]),
    "sfb200_host_free": (None, [C.c_void_p]),
    "sfb200_comm_unique_id": (C.c_int, [u8p]),
    "sfb200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p]),
    "sfb200_index_build": (C.c_int, [C.c_void_p, C.c_void_p, u64p, u32p, C.c_uint32, C.c_int]),
    "sfb200_index_stats": (C.c_int, [C.c_void_p, u64p]),
    "sfb200_index_export": (C.c_int, [C.c_void_p, u64p, u32p, u32p]),
    "sfb200_index_export_table": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sfb200_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sfb200_index_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sfb200_last_map_kernel_ms": (C.c_double, [C.c_void_p]),
    "sfb200_map_h2d_bytes": (C.c_uint64, [C.c_void_p]),
    "sfb200_map_clipped": (C.c_uint64, [C.c_void_p]),
    "sfb200_map_begin": (C.c_int, [C.c_void_p, C.POINTER(MapOpts)]),
    "sfb200_map_batch": (C.c_int, [C.c_void_p, C.c_void_p, u64p, C.c_void_p, u64p, C.c_uint64]),
    "sfb200_map_batch_fixed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64]),
    "sfb200_map_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "sfb200_map_fastq": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint64, u64p, u64p, u64p]),
    "sfb200_map_set_bias": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int32]),
    "sfb200_map_get_bias": (C.c_int, [C.c_void_p, u32p, u32p]),
    "sfb200_map_finish": (C.c_int, [C.c_void_p, u64p, u32p, u64p, u64p]),
    "sfb200_eq_export": (C.c_int, [C.c_void_p, u64p, u32p, u64p]),
    "sfb200_eq_import": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, u64p, u32p, u64p]),
    "sfb200_em_default_opts": (None, [C.POINTER(EMOpts)]),
    "sfb200_em_run": (C.c_int, [C.c_void_p, f64p, C.c_uint32, C.c_uint64, C.POINTER(EMOpts), f64p, u32p, f64p]),
    "sfb200_last_em_loop_ms": (C.c_double, [C.c_void_p]),
    "sfb200_last_em_kernel": (C.c_int, [C.c_void_p]),
    "sfb200_last_em_variant": (C.c_int, [C.c_void_p]),
    "sfb200_bias_eff_lens": (C.c_int, [C.c_void_p, C.POINTER(BiasModel), f64p, f64p, f64p, C.c_uint32, f64p]),
    "sfb200_em_run_bias": (C.c_int, [C.c_void_p, f64p, C.c_uint32, C.c_uint64, C.POINTER(EMOpts), C.POINTER(BiasModel), f64p, f64p, u32p, f64p]),
    "sfb200_bootstrap_run": (C.c_int, [C.c_void_p, f64p, C.c_uint32, C.POINTER(EMOpts), C.c_uint32, C.c_uint64, F64_ROW_CB, C.c_void_p]),
    "sfb200_bootstrap_em": (C.c_int, [C.c_void_p, f64p, C.c_uint32, u64p, C.POINTER(EMOpts), f64p, u32p]),
    "sfb200_gibbs_run": (C.c_int, [C.c_void_p, f64p, f64p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, I32_ROW_CB, C.c_void_p]),
    "sfb200_xxh64_device": (C.c_int, [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, u64p]),
    "sfb200_digamma_device": (C.c_int, [C.c_void_p, f64p, C.c_uint64, f64p]),
    "sfb200_exp_digamma_device": (C.c_int, [C.c_void_p, f64p, C.c_uint64, f64p]),
}


def build(force=False):
    """Compile sailfish_b200/csrc/*.cu for sm_100a into sailfish_b200/libsfb200.so (in-tree)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", src], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def bind_host_near_device(device=0):
    """Keep this process (and the threads it starts from now on) on the CPUs of the GPU's NUMA node, so that page-locked buffers
    allocated afterwards are node-local.  -> number of CPUs bound, 0 if nothing changed (include/sfb200.h)."""
    rc = lib().sfb200_bind_host_near_device(int(device))
    if rc < 0:
        raise Sfb200Error(rc, "sfb200_bind_host_near_device(%d)" % device)
    return rc


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def pack_reads(reads):
    """list of str/bytes -> (uint8 array, uint64 offsets[n+1])"""
    bs = [r if isinstance(r, bytes) else r.encode() for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8), off


class Context:
    """One sfb200_ctx: a CUDA device, its index, its class table and inference scratch."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sfb200_ctx_create(device, C.byref(h))
        if rc != 0:
            raise Sfb200Error(rc, "sfb200_ctx_create failed (no usable CUDA device: there is no CPU fallback)")
        self.h = h
        self.n_txp = 0
        self.map_opts = None

    def close(self):
        if getattr(self, "h", None):
            self.L.sfb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise Sfb200Error(rc, self.L.sfb200_last_error(self.h).decode(errors="replace"))

    # ---- plumbing
    def set_stream(self, cuda_stream_ptr):
        self._chk(self.L.sfb200_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self):
        self._chk(self.L.sfb200_ctx_sync(self.h))

    def launch_count(self):
        return int(self.L.sfb200_launch_count(self.h))

    def comm_init(self, n_ranks, rank, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._chk(self.L.sfb200_comm_init(self.h, n_ranks, rank, _ptr(uid, u8p)))

    @staticmethod
    def comm_unique_id():
        uid = np.zeros(128, np.uint8)
        rc = lib().sfb200_comm_unique_id(_ptr(uid, u8p))
        if rc != 0:
            raise Sfb200Error(rc, "ncclGetUniqueId")
        return uid

    # ---- index
    def index_build(self, seqs=None, k=31, seq=None, txp_off=None, txp_len=None):
        """seqs: list of transcript sequences; or seq (uint8/bytes blob) + txp_off + txp_len."""
        if seqs is not None:
            bs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
            seq = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
            txp_len = np.array([len(b) for b in bs], dtype=np.uint32)
            txp_off = np.zeros(len(bs), dtype=np.uint64)
            if len(bs) > 1:
                txp_off[1:] = np.cumsum(txp_len.astype(np.uint64))[:-1]
        if isinstance(seq, (bytes, bytearray)):
            seq = np.frombuffer(bytes(seq) + b"\0", dtype=np.uint8)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        txp_off = np.ascontiguousarray(txp_off, dtype=np.uint64)
        txp_len = np.ascontiguousarray(txp_len, dtype=np.uint32)
        self._chk(self.L.sfb200_index_build(self.h, seq.ctypes.data_as(C.c_void_p), _ptr(txp_off, u64p), _ptr(txp_len, u32p),
                                            len(txp_len), k))
        self.n_txp = len(txp_len)
        self.txp_len = txp_len
        return self.index_stats()

    def index_save(self, path):
        self._chk(self.L.sfb200_index_save(self.h, os.fsencode(path)))

    def index_load(self, path):
        self._chk(self.L.sfb200_index_load(self.h, os.fsencode(path)))
        return self.index_stats()

    def index_stats(self):
        st = np.zeros(8, np.uint64)
        self._chk(self.L.sfb200_index_stats(self.h, _ptr(st, u64p)))
        return dict(text_len=int(st[0]), n_sa=int(st[1]), n_kmers=int(st[2]), table_slots=int(st[3]), hbm_bytes=int(st[4]),
                    max_bucket=int(st[5]), k=int(st[6]), n_txp=int(st[7]))

    def index_export(self):
        st = self.index_stats()
        words = np.zeros(st["text_len"] // 32 + 2, np.uint64)
        sa_pos = np.zeros(max(st["n_sa"], 1), np.uint32)
        sa_tid = np.zeros(max(st["n_sa"], 1), np.uint32)
        self._chk(self.L.sfb200_index_export(self.h, _ptr(words, u64p), _ptr(sa_pos, u32p), _ptr(sa_tid, u32p)))
        return words, sa_pos[:st["n_sa"]], sa_tid[:st["n_sa"]]

    def index_export_table(self):
        st = self.index_stats()
        tab = np.zeros((st["table_slots"], 2), np.uint64)
        self._chk(self.L.sfb200_index_export_table(self.h, tab.ctypes.data_as(C.c_void_p)))
        return tab

    def last_map_kernel_ms(self):
        return float(self.L.sfb200_last_map_kernel_ms(self.h))

    def map_h2d_bytes(self):
        return int(self.L.sfb200_map_h2d_bytes(self.h))

    # ---- mapping
    def map_clipped(self):
        """mates longer than 256 bases (mapped by their first 256) since map_begin"""
        return int(self.L.sfb200_map_clipped(self.h))

    def map_begin(self, opts):
        self.map_opts = opts
        self._chk(self.L.sfb200_map_begin(self.h, C.byref(opts)))

    def map_batch(self, bases1, off1, bases2=None, off2=None):
        """HOST buffers (numpy uint8 + uint64 offsets); includes the H2D copy."""
        off1 = np.ascontiguousarray(off1, dtype=np.uint64)
        n = len(off1) - 1
        b1 = np.ascontiguousarray(bases1, dtype=np.uint8)
        if bases2 is not None:
            off2 = np.ascontiguousarray(off2, dtype=np.uint64)
            b2 = np.ascontiguousarray(bases2, dtype=np.uint8)
            self._chk(self.L.sfb200_map_batch(self.h, b1.ctypes.data_as(C.c_void_p), _ptr(off1, u64p),
                                              b2.ctypes.data_as(C.c_void_p), _ptr(off2, u64p), n))
        else:
            self._chk(self.L.sfb200_map_batch(self.h, b1.ctypes.data_as(C.c_void_p), _ptr(off1, u64p), None, None, n))

    def map_batch_ptr(self, p_bases1, p_off1, p_bases2, p_off2, n, device):
        """raw pointers (pinned host memory with device=False, device memory with device=True)"""
        f = self.L.sfb200_map_batch_device if device else self.L.sfb200_map_batch
        if device:
            self._chk(f(self.h, C.c_void_p(p_bases1), C.c_void_p(p_off1), C.c_void_p(p_bases2) if p_bases2 else None,
                        C.c_void_p(p_off2) if p_off2 else None, n))
        else:
            self._chk(f(self.h, C.c_void_p(p_bases1), C.cast(C.c_void_p(p_off1), u64p), C.c_void_p(p_bases2) if p_bases2 else None,
                        C.cast(C.c_void_p(p_off2), u64p) if p_off2 else None, n))

    def map_batch_fixed(self, b1, len1, b2=None, len2=0, n=None):
        """host arrays of reads of one length per mate, stored back to back (no offsets)"""
        b1 = np.ascontiguousarray(b1, dtype=np.uint8)
        if n is None:
            n = len(b1) // int(len1) if len1 else 0
        p2 = None
        if b2 is not None:
            b2 = np.ascontiguousarray(b2, dtype=np.uint8)
            p2 = b2.ctypes.data_as(C.c_void_p)
        self._chk(self.L.sfb200_map_batch_fixed(self.h, b1.ctypes.data_as(C.c_void_p), int(len1), p2, int(len2), int(n)))

    def map_batch_fixed_ptr(self, p_bases1, len1, p_bases2, len2, n):
        """raw pointers to (pinned) host memory"""
        self._chk(self.L.sfb200_map_batch_fixed(self.h, C.c_void_p(p_bases1), int(len1), C.c_void_p(p_bases2) if p_bases2 else None, int(len2), int(n)))

    def map_fastq(self, text1, text2=None, max_records=0):
        """FASTQ TEXT (bytes, starting at a record boundary) -> extracted and mapped on the device; -> (records, consumed1, consumed2)"""
        n = C.c_uint64(); c1 = C.c_uint64(); c2 = C.c_uint64()
        self._chk(self.L.sfb200_map_fastq(self.h, text1, len(text1), text2, len(text2) if text2 is not None else 0, int(max_records),
                                          C.byref(n), C.byref(c1), C.byref(c2) if text2 is not None else None))
        return n.value, c1.value, c2.value

    def map_set_bias(self, seq_bias=True, gc_bias=False, num_bias_samples=1000000):
        """collect the bias / GC samples while mapping (after map_begin, before the first batch)"""
        self._chk(self.L.sfb200_map_set_bias(self.h, int(seq_bias), int(gc_bias), int(num_bias_samples)))

    def map_get_bias(self):
        """-> (read_bias[4096], observed_gc[101]) with the initial count of 1 per bin"""
        rb = np.zeros(4096, np.uint32); og = np.zeros(101, np.uint32)
        self._chk(self.L.sfb200_map_get_bias(self.h, _ptr(rb, u32p), _ptr(og, u32p)))
        return rb, og

    def map_finish(self):
        counters = np.zeros(6, np.uint64)
        fld = np.zeros(self.map_opts.max_frag_len, np.uint32)
        E = C.c_uint64(); nnz = C.c_uint64()
        self._chk(self.L.sfb200_map_finish(self.h, _ptr(counters, u64p), _ptr(fld, u32p), C.byref(E), C.byref(nnz)))
        self.E, self.nnz = E.value, nnz.value
        return dict(counters=counters, fld=fld, n_classes=E.value, nnz=nnz.value)

    def eq_export(self):
        row_ptr = np.zeros(self.E + 1, np.uint64); labels = np.zeros(max(self.nnz, 1), np.uint32)
        counts = np.zeros(max(self.E, 1), np.uint64)
        self._chk(self.L.sfb200_eq_export(self.h, _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p)))
        return row_ptr, labels[:self.nnz], counts[:self.E]

    def eq_import(self, n_txp, row_ptr, labels, counts):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        labels = np.ascontiguousarray(labels, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self.n_txp, self.E, self.nnz = n_txp, len(counts), len(labels)
        self._chk(self.L.sfb200_eq_import(self.h, n_txp, len(counts), _ptr(row_ptr, u64p), _ptr(labels, u32p), _ptr(counts, u64p)))

    # ---- inference
    def em_run(self, eff_lens, num_mapped, opts=None):
        opts = opts or EMOpts.default()
        eff = np.ascontiguousarray(eff_lens, dtype=np.float64)
        alphas = np.zeros(len(eff), np.float64)
        iters = C.c_uint32(); mrd = C.c_double()
        self._chk(self.L.sfb200_em_run(self.h, _ptr(eff, f64p), len(eff), int(num_mapped), C.byref(opts), _ptr(alphas, f64p),
                                       C.byref(iters), C.byref(mrd)))
        return alphas, iters.value, mrd.value

    def last_em_loop_ms(self):
        return float(self.L.sfb200_last_em_loop_ms(self.h))

    def bias_eff_lens(self, mode, eff_model, eff_in, alphas, num_fwd, num_rc, read_bias, observed_gc, fld_cdf, fld_max, gc_samp=1):
        """updateEffectiveLengths on the device (sfb200_bias_eff_lens); fld_cdf = EmpiricalDistribution's float cdf table"""
        eff_model = np.ascontiguousarray(eff_model, dtype=np.float64); eff_in = np.ascontiguousarray(eff_in, dtype=np.float64)
        alphas = np.ascontiguousarray(alphas, dtype=np.float64)
        rb = np.ascontiguousarray(read_bias, dtype=np.uint32); og = np.ascontiguousarray(observed_gc, dtype=np.uint32)
        cdf = np.ascontiguousarray(fld_cdf, dtype=np.float32)
        m = BiasModel(int(mode), int(gc_samp), int(num_fwd), int(num_rc), _ptr(rb, u32p), _ptr(og, u32p),
                      cdf.ctypes.data_as(C.POINTER(C.c_float)), len(cdf), int(fld_max))
        out = np.zeros(len(eff_in), np.float64)
        self._chk(self.L.sfb200_bias_eff_lens(self.h, C.byref(m), _ptr(eff_model, f64p), _ptr(eff_in, f64p), _ptr(alphas, f64p), len(eff_in),
                                              _ptr(out, f64p)))
        return out

    @staticmethod
    def _bias_model(mode, num_fwd, num_rc, read_bias, observed_gc, fld_cdf, fld_max, gc_samp):
        rb = np.ascontiguousarray(read_bias, dtype=np.uint32); og = np.ascontiguousarray(observed_gc, dtype=np.uint32)
        cdf = np.ascontiguousarray(fld_cdf, dtype=np.float32)
        m = BiasModel(int(mode), int(gc_samp), int(num_fwd), int(num_rc), _ptr(rb, u32p), _ptr(og, u32p),
                      cdf.ctypes.data_as(C.POINTER(C.c_float)), len(cdf), int(fld_max))
        return m, (rb, og, cdf)                                    # the arrays must outlive the call

    def em_run_bias(self, mode, eff_lens, num_mapped, num_fwd, num_rc, read_bias, observed_gc, fld_cdf, fld_max, gc_samp=1, opts=None):
        """optimize() with bias (mode 1) / GC (mode 2) correction (sfb200_em_run_bias) -> (alphas, final eff lens, iters, max_rel_diff)"""
        opts = opts or EMOpts.default()
        eff = np.ascontiguousarray(eff_lens, dtype=np.float64)
        m, keep = self._bias_model(mode, num_fwd, num_rc, read_bias, observed_gc, fld_cdf, fld_max, gc_samp)
        alphas = np.zeros(len(eff), np.float64); eff_out = np.zeros(len(eff), np.float64)
        iters = C.c_uint32(); mrd = C.c_double()
        self._chk(self.L.sfb200_em_run_bias(self.h, _ptr(eff, f64p), len(eff), int(num_mapped), C.byref(opts), C.byref(m), _ptr(alphas, f64p),
                                            _ptr(eff_out, f64p), C.byref(iters), C.byref(mrd)))
        del keep
        return alphas, eff_out, iters.value, mrd.value

    def last_em_variant(self):
        """bit 0: k_em_dense streamed its counts from global memory; bit 1: lagged stopping rule"""
        return int(self.L.sfb200_last_em_variant(self.h))

    def last_em_kernel(self):
        """0 k_em_persistent, 1 k_em_part, 2 k_em_gather, 3 one launch per phase, 4 k_em_dense, 5 k_em_dense with a pool loop (hybrid)"""
        return int(self.L.sfb200_last_em_kernel(self.h))

    def bootstrap_em(self, eff_lens, samp_counts, opts=None):
        opts = opts or EMOpts.default()
        eff = np.ascontiguousarray(eff_lens, dtype=np.float64)
        sc = np.ascontiguousarray(samp_counts, dtype=np.uint64)
        alphas = np.zeros(len(eff), np.float64)
        iters = C.c_uint32()
        self._chk(self.L.sfb200_bootstrap_em(self.h, _ptr(eff, f64p), len(eff), _ptr(sc, u64p), C.byref(opts), _ptr(alphas, f64p),
                                             C.byref(iters)))
        return alphas, iters.value

    def bootstrap_run(self, eff_lens, n_boot, seed=1, opts=None):
        opts = opts or EMOpts.default()
        eff = np.ascontiguousarray(eff_lens, dtype=np.float64)
        rows = []
        cb = F64_ROW_CB(lambda u, p, n: (rows.append(np.ctypeslib.as_array(p, shape=(n,)).copy()), 0)[1])
        self._chk(self.L.sfb200_bootstrap_run(self.h, _ptr(eff, f64p), len(eff), C.byref(opts), n_boot, seed, cb, None))
        return np.array(rows)

    def gibbs_run(self, eff_lens, masses, num_mapped, n_samples, seed=1):
        eff = np.ascontiguousarray(eff_lens, dtype=np.float64)
        masses = np.ascontiguousarray(masses, dtype=np.float64)
        rows = []
        cb = I32_ROW_CB(lambda u, p, n: (rows.append(np.ctypeslib.as_array(p, shape=(n,)).copy()), 0)[1])
        self._chk(self.L.sfb200_gibbs_run(self.h, _ptr(eff, f64p), _ptr(masses, f64p), len(eff), int(num_mapped), n_samples, seed,
                                          cb, None))
        return np.array(rows)

    # ---- known-answer hooks
    def xxh64(self, msgs, seed=0):
        bs = [bytes(m) for m in msgs]
        off = np.zeros(len(bs) + 1, np.uint64)
        off[1:] = np.cumsum([len(b) for b in bs])
        data = np.frombuffer(b"".join(bs) + b"\0\0\0\0", dtype=np.uint8)
        out = np.zeros(len(bs), np.uint64)
        self._chk(self.L.sfb200_xxh64_device(self.h, _ptr(data, u8p), _ptr(off, u64p), len(bs), seed, _ptr(out, u64p)))
        return out

    def digamma(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(len(x), np.float64)
        self._chk(self.L.sfb200_digamma_device(self.h, _ptr(x, f64p), len(x), _ptr(out, f64p)))
        return out

    def exp_digamma(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(len(x), np.float64)
        self._chk(self.L.sfb200_exp_digamma_device(self.h, _ptr(x, f64p), len(x), _ptr(out, f64p)))
        return out


def tpm(alphas, eff_lens, num_mapped):
    """TPM column of quant.sf (reference src/GZipWriter.cpp:216-241)."""
    npm = np.asarray(alphas, np.float64) / float(num_mapped)
    tfrac = npm / np.asarray(eff_lens, np.float64)
    return tfrac / tfrac.sum() * 1e6
