"""`sailfish quant`-shaped driver (sailfish_b200/quant.py): format helpers on CPU; on a GPU, BASELINE config 1 -- the bundled
sample data (15 transcripts, 10 000 read pairs x 50 nt, -l IU) end to end from FASTA/FASTQ files to quant.sf."""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import split_seqs
from sailfish_b200 import quant


def test_library_format_parsing_matches_oracle():
    from oracle import pyoracle as O
    for name in ["IU", "ISF", "ISR", "OU", "OSF", "OSR", "MU", "MSF", "MSR", "U", "SF", "SR", "iu"]:
        assert quant.parse_library_format(name) == O.parse_libtype(name)
    with pytest.raises(ValueError):
        quant.parse_library_format("XYZ")


def test_percent_g_formatting():
    # cppformat's `{}` for doubles is printf("%g") (reference include/spdlog/details/format.h:2895-2912)
    assert quant.fmt_g(0.0) == "0" and quant.fmt_g(1234567.0) == "1.23457e+06" and quant.fmt_g(0.000123456789) == "0.000123457"
    assert quant.fmt_g(100.0) == "100" and quant.fmt_g(33.3333333) == "33.3333"


def test_fragment_length_tables_for_the_bias_model():
    """efflen.empirical_cdf == EmpiricalDistribution's float cdf table (the oracle's restatement, pinned to the reference's class in
    tests/test_oracle_bias.py), bit for bit; efflen.normal_frag_length_counts == getNormalFragLengthCounts"""
    from oracle import pyoracle as O
    from sailfish_b200 import efflen
    rng = np.random.default_rng(1)
    x = np.arange(1000)
    sparse = np.zeros(1000, np.uint32); sparse[rng.integers(0, 1000, 40)] = rng.integers(1, 50, 40)
    cases = [np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32), efflen.normal_frag_length_counts(),
             rng.integers(0, 1000, size=1000).astype(np.uint32), np.full(300, 7, np.uint32), sparse, np.array([5], np.uint32)]
    for c in cases:
        got, mx = efflen.empirical_cdf(c)
        want, mx_o = O.fld_cdf(c)
        assert mx == mx_o and got.tobytes() == np.asarray(want, np.float32).tobytes()
    nc = efflen.normal_frag_length_counts(1000, 10000, 200.0, 80.0)
    assert nc.dtype == np.uint32 and len(nc) == 1000 and int(nc[200]) == int(nc.max()) and abs(int(nc.sum()) - 10000) < 20
    dens = np.exp(-0.5 * ((x - 200.0) / 80.0) ** 2) / 80.0
    assert nc.tolist() == np.floor(dens * 10000 / dens.sum() + 0.5).astype(np.uint32).tolist()


def test_aux_vectors(tmp_path):
    """aux/fld.gz, observed_*.gz, expected_*.gz in GZipWriter::writeMeta's layout (src/GZipWriter.cpp:139-161)"""
    from sailfish_b200 import efflen
    x = np.arange(1000)
    fld = np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)
    rb = np.arange(1, 4097, dtype=np.uint32); og = np.arange(1, 102, dtype=np.uint32)
    quant.write_aux_vectors(str(tmp_path), fld, rb, og)
    real = np.frombuffer(gzip.open(tmp_path / "fld.gz").read(), dtype=np.int32)
    assert len(real) == 1000 and int(real.sum()) == 10000 and abs(float((real * x).sum()) / 10000 - 190.0) < 3
    assert np.frombuffer(gzip.open(tmp_path / "observed_bias.gz").read(), dtype=np.int32).tolist() == rb.tolist()
    assert np.frombuffer(gzip.open(tmp_path / "observed_gc.gz").read(), dtype=np.int32).tolist() == og.tolist()
    assert (np.frombuffer(gzip.open(tmp_path / "expected_bias.gz").read(), dtype=np.float64) == np.ones(4096)).all()
    assert (np.frombuffer(gzip.open(tmp_path / "expected_gc.gz").read(), dtype=np.float64) == np.ones(101)).all()
    (tmp_path / "z").mkdir()
    quant.write_aux_vectors(str(tmp_path / "z"), np.zeros(1000, np.uint32), rb, og)               # nothing observed: no draws, no crash
    assert not np.frombuffer(gzip.open(tmp_path / "z" / "fld.gz").read(), dtype=np.int32).any()
    quant.write_aux_vectors(str(tmp_path / "z"), None, rb, og)


class _StubContext:
    """Test double for capi.Context: records the calls of the driver and returns canned device results, so that the HOST flow of
    sailfish_b200.quant (option checks, effective lengths, FLD hand-over, output files) runs in the CPU suite.  Not a CPU fallback:
    it computes nothing."""
    calls = []

    def __init__(self, device=0):
        type(self).calls = []

    def _rec(self, name, *a, **k):
        type(self).calls.append((name, a, k))

    def index_build(self, seqs=None, k=31, **kw):
        self._rec("index_build", len(seqs), k); self.T = len(seqs)

    def map_begin(self, opts):
        self._rec("map_begin"); self.map_opts = opts

    def map_set_bias(self, seq_bias, gc_bias, n):
        self._rec("map_set_bias", bool(seq_bias), bool(gc_bias), int(n))

    def map_batch(self, b1, o1, b2=None, o2=None):
        self._rec("map_batch", len(o1) - 1, b2 is not None)

    def map_fastq(self, t1, t2=None, max_records=0):
        def recs(t):
            return t.count(b"\n") // 4

        def consumed(t, n):
            pos = -1
            for _ in range(4 * n):
                pos = t.index(b"\n", pos + 1)
            return pos + 1
        n = min(recs(t1), recs(t2)) if t2 is not None else recs(t1)
        assert n == 0 or t1[:1] == b"@"                      # every block must start at a record boundary
        self._rec("map_fastq", n, t2 is not None)
        return n, consumed(t1, n), consumed(t2, n) if t2 is not None else 0

    def map_clipped(self):
        return 0

    def map_finish(self):
        fld = np.zeros(self.map_opts.max_frag_len, np.uint32); fld[180:220] = 300          # 12000 sampled fragment lengths
        return dict(counters=np.array([4, 3, 5, 4, 2, 1], np.uint64), fld=fld, n_classes=2, nnz=3)

    def map_get_bias(self):
        return np.arange(1, 4097, dtype=np.uint32), np.arange(1, 102, dtype=np.uint32)

    def eq_export(self):
        return np.array([0, 1, 3], np.uint64), np.array([0, 0, 1], np.uint32), np.array([2, 1], np.uint64)

    def em_run(self, eff, num_mapped, opts=None):
        self._rec("em_run", np.array(eff), num_mapped)
        return np.array([2.0, 1.0] + [0.0] * (self.T - 2)), 51, 0.001

    def em_run_bias(self, mode, eff, num_mapped, nf, nr, rb, og, cdf, fld_max, gc_samp=1, opts=None):
        self._rec("em_run_bias", mode, nf, nr, len(cdf), fld_max, gc_samp)
        return np.array([2.0, 1.0] + [0.0] * (self.T - 2)), np.array(eff) * 0.5, 77, 0.002

    def bootstrap_run(self, eff, n, opts=None):
        return np.ones((n, self.T))

    def gibbs_run(self, eff, masses, num_mapped, n):
        return np.ones((n, self.T), np.int32)

    def close(self):
        pass


def test_driver_host_flow_with_stub_device(tmp_path, monkeypatch):
    monkeypatch.setattr(quant.capi, "Context", _StubContext)
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 100 + "\n>t1\n" + "GGCA" * 150 + "\n>t2\n" + "TTGA" * 60 + "\n")
    for tag in "12":
        (tmp_path / ("r%s.fq" % tag)).write_text("".join("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 10, "I" * 40) for i in range(4)))
    base = ["-t", str(fa), "-l", "IU", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq")]
    # plain run: smoothed effective lengths from the observed FLD, every aux file, no bias calls
    out = tmp_path / "o1"
    quant.main(base + ["-o", str(out), "--dumpEq", "--numBootstraps", "2"])
    names = [c[0] for c in _StubContext.calls]
    assert names == ["index_build", "map_begin", "map_batch", "em_run"]
    rows = [l.split("\t") for l in open(out / "quant.sf").read().strip().split("\n")[1:]]
    assert [r[0] for r in rows] == ["t0", "t1", "t2"] and [r[4] for r in rows] == ["2", "1", "0"]
    np.testing.assert_allclose([float(r[2]) for r in rows], [400 - 199.5 + 1, 600 - 199.5 + 1, 240 - 199.5 + 1], rtol=1e-6)
    meta = json.load(open(out / "aux" / "meta_info.json"))
    assert meta["frag_dist_length"] == 999 and meta["bias_correct"] is False and meta["num_bias_bins"] == 4096 and meta["samp_type"] == "bootstrap"
    assert meta["num_processed"] == 4 and meta["num_mapped"] == 3 and "start_time" in meta
    real = np.frombuffer(gzip.open(out / "aux" / "fld.gz").read(), dtype=np.int32)
    assert len(real) == 1000 and real.sum() == 10000 and real[:180].sum() == 0 and real[220:].sum() == 0
    assert (np.frombuffer(gzip.open(out / "aux" / "observed_bias.gz").read(), dtype=np.int32) == 1).all()
    assert len(np.frombuffer(gzip.open(out / "aux" / "bootstrap" / "bootstraps.gz").read(), dtype=np.float64)) == 6
    # --unsmoothedFLD: sum_l pdf(l) (len - l + 1) with a flat pdf on 180..219
    out = tmp_path / "o2"
    quant.main(base + ["-o", str(out), "--unsmoothedFLD"])
    eff = [c for c in _StubContext.calls if c[0] == "em_run"][0][1][0]
    # EmpiricalDistribution drops the bin at which the cumulative mass passes 1 - 1e-6 (here the last one, 219): mean of 180..218
    np.testing.assert_allclose(eff, [400 - 199.0 + 1, 600 - 199.0 + 1, 240 - 199.0 + 1], rtol=1e-5)
    # --gcBiasCorrect: samples requested before the first batch, the optimizer variant gets the FLD table, corrected lengths are reported
    out = tmp_path / "o3"
    quant.main(base + ["-o", str(out), "--gcBiasCorrect", "--gcSpeedSamp", "3", "--numBiasSamples", "1234"])
    calls = _StubContext.calls
    assert [c[0] for c in calls] == ["index_build", "map_begin", "map_set_bias", "map_batch", "em_run_bias"]
    assert calls[2][1] == (False, True, 1234)
    mode, nf, nr, n_cdf, fld_max, gc_samp = calls[4][1]
    assert (mode, nf, nr, fld_max, gc_samp) == (2, 2, 1, 999, 3) and n_cdf == 219
    rows = [l.split("\t") for l in open(out / "quant.sf").read().strip().split("\n")[1:]]
    np.testing.assert_allclose([float(r[2]) for r in rows], [0.5 * (400 - 199.5 + 1), 0.5 * (600 - 199.5 + 1), 0.5 * (240 - 199.5 + 1)], rtol=1e-5)
    assert json.load(open(out / "aux" / "meta_info.json"))["bias_correct"] is False                          # opts.biasCorrect only
    assert np.frombuffer(gzip.open(out / "aux" / "observed_gc.gz").read(), dtype=np.int32).tolist() == list(range(1, 102))
    # --deviceParse: raw text block-wise, the unconsumed tail carried over; mate 2 has longer records and no final newline
    (tmp_path / "d1.fq").write_text("".join("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 10, "I" * 40) for i in range(50)))
    (tmp_path / "d2.fq").write_text("".join("@r%d with a long description\n%s\n+\n%s\n" % (i, "ACGT" * 20, "@" * 80) for i in range(50))[:-1])
    for blk in ("100000", "1000", "150"):
        quant.main(["-t", str(fa), "-l", "IU", "-1", str(tmp_path / "d1.fq"), "-2", str(tmp_path / "d2.fq"), "-o", str(tmp_path / ("od" + blk)),
                    "--deviceParse", "--blockBytes", blk])
        calls = [c for c in _StubContext.calls if c[0] == "map_fastq"]
        assert sum(c[1][0] for c in calls) == 50 and all(c[1][1] for c in calls) and "map_batch" not in [c[0] for c in _StubContext.calls]
        assert blk == "100000" or len(calls) > 5
    with pytest.raises(ValueError):
        quant.main(["-t", str(fa), "-l", "IU", "-1", str(tmp_path / "d1.fq"), "-2", str(tmp_path / "r2.fq"), "-o", str(tmp_path / "odx"), "--deviceParse"])
    # option checks (SailfishQuantify.cpp:1293-1309)
    with pytest.raises(ValueError):
        quant.main(base + ["-o", str(tmp_path / "o4"), "--biasCorrect", "--gcBiasCorrect"])
    quant.main(["-t", str(fa), "-l", "U", "-r", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o5"), "--gcBiasCorrect"])
    assert "map_set_bias" not in [c[0] for c in _StubContext.calls]                                         # switched off for single-end reads


def test_fastx_readers(tmp_path):
    fa = tmp_path / "t.fa"; fa.write_text(">a desc\nACGT\nAC\n>b\nGG\n")
    assert quant.read_fasta(str(fa)) == (["a", "b"], ["ACGTAC", "GG"])
    fq = tmp_path / "r.fq"; fq.write_text("@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\nIIII\n@r3\nTT\n+\nII\n")
    assert list(quant.read_fastx_batches(str(fq), 2)) == [["ACGT", "GGCC"], ["TT"]]
    with gzip.open(tmp_path / "r.fa.gz", "wt") as f:
        f.write(">x\nAC\nGT\n>y\nTTT\n")
    assert list(quant.read_fastx_batches(str(tmp_path / "r.fa.gz"), 10)) == [["ACGT", "TTT"]]


@pytest.mark.gpu
def test_sample_data_end_to_end(sample_data, tmp_path):
    d = sample_data
    seqs = split_seqs(d["txp_seq"], d["txp_len"])
    names = [str(n) for n in d["names"]]
    fa = tmp_path / "transcripts.fasta"
    with open(fa, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n%s\n" % (n, s.decode()))
    for tag, reads, off in (("1", d["reads1"], d["off1"]), ("2", d["reads2"], d["off2"])):
        with open(tmp_path / ("reads_%s.fastq" % tag), "w") as f:
            for i in range(len(off) - 1):
                s = reads[int(off[i]):int(off[i + 1])].tobytes().decode()
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))
    out = tmp_path / "sample_quant"
    quant.main(["-t", str(fa), "-l", "IU", "-1", str(tmp_path / "reads_1.fastq"), "-2", str(tmp_path / "reads_2.fastq"),
                "-o", str(out), "--dumpEq", "--numBootstraps", "3"])
    assert (out / "quant.sf").exists()                       # what the reference's own smoke test checks (cmake/SimpleTest.cmake:34-38)
    lines = open(out / "quant.sf").read().strip().split("\n")
    assert lines[0] == "Name\tLength\tEffectiveLength\tTPM\tNumReads" and len(lines) == 1 + len(names)
    rows = [l.split("\t") for l in lines[1:]]
    assert [r[0] for r in rows] == names and [int(r[1]) for r in rows] == [int(x) for x in d["txp_len"]]
    num_reads = np.array([float(r[4]) for r in rows]); tpm = np.array([float(r[3]) for r in rows])
    np.testing.assert_allclose(num_reads, d["ref_est_vb0"], rtol=1.2e-4, atol=1e-6)     # 1e-4 parity + 6 significant digits
    np.testing.assert_allclose([float(r[2]) for r in rows], d["eff"], rtol=1e-5)
    assert abs(tpm.sum() - 1e6) < 50
    meta = json.load(open(out / "aux" / "meta_info.json"))
    assert meta["num_processed"] == 10000 and meta["num_mapped"] == int(d["num_mapped"]) and meta["samp_type"] == "bootstrap"
    eq = open(out / "aux" / "eq_classes.txt").read().split("\n")
    assert int(eq[0]) == len(names) and int(eq[1]) == len(d["counts"])
    first = eq[2 + len(names)].split("\t")
    assert int(first[0]) == len(first) - 2
    boots = np.frombuffer(gzip.open(out / "aux" / "bootstrap" / "bootstraps.gz").read(), dtype=np.float64).reshape(3, len(names))
    np.testing.assert_allclose(boots.sum(axis=1), float(d["num_mapped"]), rtol=1e-9)
