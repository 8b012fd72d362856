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
