"""The C++ host driver sfb200-quant (sailfish_b200/host/sfb200_quant.cpp + fastx_reader.hpp): the FASTA/FASTQ ingestion is
checked on CPU against a plain Python parse (ragged reads, CRLF, no trailing newline, gzip, block and batch boundaries);
on a GPU the whole command reproduces the reference optimizer's estimates on the bundled sample data (BASELINE config 1)."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import split_seqs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sailfish_b200", "bin", "sfb200-quant")


def build_exe():
    from sailfish_b200 import capi
    capi.lib()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sailfish_b200", "host")])
    return EXE


def fnv(seqs1, seqs2):
    M = (1 << 64) - 1

    def h(vals):
        x = 1469598103934665603
        for v in vals:
            x = ((x ^ v) * 1099511628211) & M
        return "%016x" % x
    return [h("".join(seqs1).encode()), h("".join(seqs2).encode()), h(len(s) & 0xFFFF for s in seqs1), h(len(s) & 0xFFFF for s in seqs2)]


def rand_reads(rng, n, lo, hi):
    return ["".join(rng.choice(list("ACGTN"), size=int(rng.integers(lo, hi + 1)))) for _ in range(n)]


def write_fastq(path, seqs, eol="\n", final_newline=True, gz=False):
    txt = eol.join("@r%d some description%s%s%s+%s%s" % (i, eol, s, eol, eol, "I" * len(s)) for i, s in enumerate(seqs))
    if final_newline:
        txt += eol
    (gzip.open(path, "wt", newline="") if gz else open(path, "w", newline="")).write(txt)


def parse_only(args):
    out = subprocess.check_output([build_exe(), "--parseOnly"] + args)
    return json.loads(out)


@pytest.mark.parametrize("eol,final_newline,gz", [("\n", True, False), ("\r\n", True, False), ("\n", False, False), ("\n", True, True)])
def test_fastq_ingestion_single(tmp_path, eol, final_newline, gz):
    rng = np.random.default_rng(3)
    seqs = rand_reads(rng, 5000, 1, 120) + ["A"] + rand_reads(rng, 10, 300, 400)
    p = str(tmp_path / ("r.fq.gz" if gz else "r.fq"))
    write_fastq(p, seqs, eol, final_newline, gz)
    for block, batch in ((64, 7), (4096, 1000), (1 << 20, 1 << 20)):
        got = parse_only(["-r", p, "--blockBytes", str(block), "--batchReads", str(batch), "-p", "3"])
        assert got["records"] == len(seqs) and got["bases1"] == sum(len(s) for s in seqs) and got["bases2"] == 0
        assert got["fnv1a"] == fnv(seqs, [])


def test_fastq_ingestion_paired_and_multiple_files(tmp_path):
    rng = np.random.default_rng(4)
    a1, a2 = rand_reads(rng, 3000, 30, 80), rand_reads(rng, 3000, 30, 80)
    b1, b2 = rand_reads(rng, 1234, 50, 50), rand_reads(rng, 1234, 50, 50)
    for name, s in (("a1", a1), ("a2", a2), ("b1", b1), ("b2", b2)):
        write_fastq(str(tmp_path / (name + ".fq")), s)
    got = parse_only(["-1", str(tmp_path / "a1.fq"), str(tmp_path / "b1.fq"), "-2", str(tmp_path / "a2.fq"), str(tmp_path / "b2.fq"),
                      "--blockBytes", "5000", "--batchReads", "999"])
    assert got["records"] == 4234
    assert got["bases1"] == sum(map(len, a1 + b1)) and got["bases2"] == sum(map(len, a2 + b2))
    assert got["fnv1a"] == fnv(a1 + b1, a2 + b2)
    # mate files of different length are an error, as in the reference's paired parser
    write_fastq(str(tmp_path / "short.fq"), a2[:-1])
    r = subprocess.run([EXE, "--parseOnly", "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "short.fq")], capture_output=True)
    assert r.returncode != 0 and b"different numbers of reads" in r.stderr
    # a malformed second mate file is reported, not a crash of the helper thread
    (tmp_path / "bad2.fq").write_text("@r0\nACGT\nIIII\n+\n" * 50)
    r = subprocess.run([EXE, "--parseOnly", "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "bad2.fq")], capture_output=True)
    assert r.returncode == 1 and b"malformed FASTQ record" in r.stderr


def test_fasta_reads_and_errors(tmp_path):
    seqs = ["ACGTACGT", "GG", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "C"]
    p = tmp_path / "r.fa"
    p.write_text("".join(">r%d\n%s\n" % (i, "\n".join(s[j:j + 5] for j in range(0, len(s), 5))) for i, s in enumerate(seqs)))
    for block in (16, 64, 1 << 20):
        got = parse_only(["-r", str(p), "--blockBytes", str(block)])
        assert got["records"] == 4 and got["fnv1a"] == fnv(seqs, [])
    bad = tmp_path / "bad.fq"
    bad.write_text("@r0\nACGT\n+\nIIII\n@r1\nAC\n")
    r = subprocess.run([EXE, "--parseOnly", "-r", str(bad)], capture_output=True)
    assert r.returncode != 0 and b"truncated" in r.stderr
    r = subprocess.run([EXE, "-t", "x.fa", "-l", "XYZ", "-r", str(p), "-o", str(tmp_path / "o")], capture_output=True)
    assert r.returncode == 2 and b"unknown library type" in r.stderr


def test_index_command(tmp_path):
    """`index` (reference src/SailfishIndexer.cpp:66-237): odd k, versionInfo.json with the reference's two fields, no rebuild
    without --force; the directory holds names, lengths and sequence (the device structures are rebuilt at load time)"""
    import struct
    fa = tmp_path / "t.fa"; fa.write_text(">t0 gene=x\nACGTAC\nGTAC\n>t1\nNNGGCC\n")
    idx = tmp_path / "idx"
    r = subprocess.run([build_exe(), "index", "-t", str(fa), "-o", str(idx), "-k", "30"], capture_output=True)
    assert r.returncode == 1 and b"should be odd" in r.stderr
    subprocess.check_call([EXE, "index", "-t", str(fa), "-o", str(idx), "-k", "21"])
    assert json.load(open(idx / "versionInfo.json")) == {"indexVersion": 2, "kmerLength": 21}
    assert json.load(open(idx / "header.json"))["NumTranscripts"] == 2
    raw = open(idx / "seq.bin", "rb").read()
    assert struct.unpack("<Q", raw[:8])[0] == 16 and raw[8:] == b"ACGTACGTACNNGGCC"
    info = open(idx / "txpInfo.bin", "rb").read()
    assert info[8:14] == struct.pack("<I", 2) + b"t0" and info[-8:] == struct.pack("<II", 10, 6)
    r = subprocess.run([EXE, "index", "-t", str(fa), "-o", str(idx)], capture_output=True)
    assert r.returncode == 0 and b"up-to-date" in r.stderr


def gene_agg_reference(lines, t2g):
    """aggregateEstimatesToGeneLevel restated independently (reference src/SailfishUtils.cpp:929-1041), including the running-sum
    normaliser of :1011-1016; genes in order of first appearance"""
    out, order, recs = [lines[0]], [], {}
    for l in lines[1:]:
        t = l.split()
        g = t2g.get(t[0], t[0])
        if g not in recs:
            recs[g] = []; order.append(g)
        recs[g].append((int(t[1]), float(t[2]), [float(x) for x in t[3:]]))
    for g in order:
        rs = recs[g]
        vals = [0.0] * len(rs[0][2]); total = 0.0
        for _, _, ev in rs:
            vals = [a + b for a, b in zip(vals, ev)]
            total += vals[0]
        if total > 5e-324:
            gl = sum(r[0] * (r[2][0] / total) for r in rs); ge = sum(r[1] * (r[2][0] / total) for r in rs)
        else:
            gl = sum(r[0] / len(rs) for r in rs); ge = sum(r[1] / len(rs) for r in rs)
        out.append("\t".join([g, "%g" % gl, "%g" % ge] + ["%g" % v for v in vals]))
    return out


def test_gene_level_aggregation(tmp_path):
    rng = np.random.default_rng(5)
    names = ["t%03d" % i for i in range(40)]
    lines = ["Name\tLength\tEffectiveLength\tTPM\tNumReads"]
    for i, n in enumerate(names):
        tpm = 0.0 if i % 7 == 0 or 8 <= i < 12 else float("%g" % rng.uniform(0, 50000))
        lines.append("%s\t%d\t%g\t%g\t%g" % (n, rng.integers(200, 5000), rng.uniform(50, 4000), tpm, tpm * 0.37))
    q = tmp_path / "quant.sf"; q.write_text("\n".join(lines) + "\n")
    t2g = {n: "g%02d" % (i // 4) for i, n in enumerate(names) if i < 36}          # the last four transcripts are not in the map
    simple = tmp_path / "t2g.tsv"; simple.write_text("".join("%s\t%s\n" % kv for kv in t2g.items()))
    r = subprocess.run([build_exe(), "genes", "-g", str(simple), "-q", str(q)], capture_output=True)
    assert r.returncode == 0 and b"36 transcripts mapping to 9 genes" in r.stderr and b"4 transcripts" in r.stderr
    assert open(tmp_path / "quant.genes.sf").read().strip().split("\n") == gene_agg_reference(lines, t2g)
    # the same map as GTF (exon lines repeat the transcript; gene_name as an alternative key)
    gtf = tmp_path / "ann.gtf"
    with open(gtf, "w") as f:
        f.write("# comment\n")
        for n, g in t2g.items():
            for feat in ("transcript", "exon", "exon"):
                f.write('chr1\tsrc\t%s\t1\t100\t.\t+\t.\tgene_id "%s"; transcript_id "%s"; gene_name "N%s";\n' % (feat, g, n, g))
    os.remove(tmp_path / "quant.genes.sf")
    subprocess.check_call([EXE, "genes", "-g", str(gtf), "-q", str(q)], stderr=subprocess.DEVNULL)
    assert open(tmp_path / "quant.genes.sf").read().strip().split("\n") == gene_agg_reference(lines, t2g)
    subprocess.check_call([EXE, "genes", "-g", str(gtf), "-q", str(q), "--txpAggregationKey", "gene_name"], stderr=subprocess.DEVNULL)
    assert open(tmp_path / "quant.genes.sf").read().strip().split("\n") == gene_agg_reference(lines, {k: "N" + v for k, v in t2g.items()})
    # a missing map stops `quant` before any work, as in the reference
    r = subprocess.run([EXE, "quant", "-t", "x.fa", "-l", "U", "-r", "y.fq", "-o", str(tmp_path / "o"), "-g", str(tmp_path / "nope.gtf")], capture_output=True)
    assert r.returncode == 1 and b"Could not find transcript <=> gene map file" in r.stderr


@pytest.mark.parametrize("mode,flags,kw", [(0, [], {}), (2, ["--unsmoothedFLD"], {"unsmoothed": True}),
                                           (1, ["--noEffectiveLengthCorrection"], {"no_correction": True})])
def test_effective_lengths_equal_oracle(tmp_path, mode, flags, kw):
    """SURVEY 8a row A9 on the host: `sfb200-quant efflens` (the function `quant` calls) and sailfish_b200/efflen.py against the
    oracle's restatement of the quasiMapReads tail -- smoothed, --unsmoothedFLD (EmpiricalDistribution's float pdf) and direct"""
    from oracle import pyoracle as O
    from sailfish_b200 import efflen
    rng = np.random.default_rng(4)
    x = np.arange(1000)
    lens = np.concatenate([rng.integers(20, 5000, size=400), [1, 5, 150, 189, 190, 191, 250, 999, 1000, 1001]]).astype(np.uint32)
    sparse = np.zeros(1000, np.uint32); sparse[rng.integers(100, 400, 60)] += rng.integers(100, 500, 60).astype(np.uint32)
    cases = [np.round(30000 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32), rng.integers(0, 100, size=1000).astype(np.uint32), sparse,
             np.round(30 * np.exp(-0.5 * ((x - 190) / 30.0) ** 2)).astype(np.uint32)]           # the last: too few samples -> prior normal
    np.savetxt(tmp_path / "lens.txt", lens, fmt="%d")
    for fld in cases:
        np.savetxt(tmp_path / "fld.txt", fld, fmt="%d")
        for single in (False, True):
            want = O.eff_lens(lens, fld, mode=mode, single_end=single)
            got_py = efflen.effective_lengths(lens, fld, single_end=single, **kw)
            out = subprocess.check_output([build_exe(), "efflens", "--lensFile", str(tmp_path / "lens.txt"), "--fldFile", str(tmp_path / "fld.txt")]
                                          + flags + (["--singleEnd"] if single else [])).decode().split()
            assert got_py.tolist() == want.tolist()
            assert [float(v) for v in out] == want.tolist()
    # aux/fld.gz as `quant` writes it: 10000 draws from the fragment length pdf as int32 counts per length
    subprocess.check_call([build_exe(), "efflens", "--lensFile", str(tmp_path / "lens.txt"), "--fldFile", str(tmp_path / "fld.txt"), "-o", str(tmp_path / "aux")],
                          stdout=subprocess.DEVNULL)
    fld_real = np.frombuffer(gzip.open(tmp_path / "aux" / "fld.gz").read(), dtype=np.int32)
    assert len(fld_real) == 1000 and int(fld_real.sum()) == 10000 and abs(float((fld_real * np.arange(1000)).sum()) / 10000 - 200.0) < 5     # prior N(200, 80)
    np.savetxt(tmp_path / "fld.txt", np.zeros(1000, np.uint32), fmt="%d")
    subprocess.check_call([build_exe(), "efflens", "--lensFile", str(tmp_path / "lens.txt"), "--fldFile", str(tmp_path / "fld.txt"), "-o", str(tmp_path / "aux0"),
                           "--numFragSamples", "0"], stdout=subprocess.DEVNULL)
    assert not np.frombuffer(gzip.open(tmp_path / "aux0" / "fld.gz").read(), dtype=np.int32).any()       # no observations: nothing to draw from


def test_driver_error_paths_with_stub_device(tmp_path):
    """a device error while mapping must end the run (the producer thread used to wait for queue room for ever), and an
    output directory that cannot be created must fail before any work is done; nested output directories are created."""
    exe = str(tmp_path / "sfb200-quant-stub")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "sailfish_b200", "host", "sfb200_quant.cpp"),
                           os.path.join(ROOT, "tests", "stub_sfb200.cpp"), "-lz"])
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 100 + "\n>t1\n" + "GGCA" * 150 + "\n")
    fq = tmp_path / "r.fq"; fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 10, "I" * 40) for i in range(400)))
    log = tmp_path / "log.txt"
    base = [exe, "quant", "-t", str(fa), "-l", "U", "-r", str(fq), "--batchReads", "10"]
    # 40 batches queued behind a failing device call: the process must exit (non-zero), not hang
    r = subprocess.run(base + ["-o", str(tmp_path / "o_fail")], env=dict(os.environ, SFB200_STUB_LOG=str(log), SFB200_STUB_FAIL_MAP="1"),
                       capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "map_batch FAILS" in log.read_text()
    # an output path below a regular file: refused before the index is built
    log.unlink()
    r = subprocess.run(base + ["-o", str(fq) + "/sub/out"], env=dict(os.environ, SFB200_STUB_LOG=str(log)), capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "cannot create directory" in (r.stderr + r.stdout)
    assert not log.exists() or "index_build" not in log.read_text()
    # nested output directories are created
    r = subprocess.run(base + ["-o", str(tmp_path / "a" / "b" / "c")], env=dict(os.environ, SFB200_STUB_LOG=str(log)), capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "a" / "b" / "c" / "quant.sf").exists() and (tmp_path / "a" / "b" / "c" / "aux" / "meta_info.json").exists()


def test_driver_device_index_file_with_stub_device(tmp_path):
    """`index --saveDeviceIndex` writes the device structures next to the sequence; `quant -i <dir>` then loads them instead of building"""
    exe = str(tmp_path / "sfb200-quant-stub")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "sailfish_b200", "host", "sfb200_quant.cpp"),
                           os.path.join(ROOT, "tests", "stub_sfb200.cpp"), "-lz"])
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 100 + "\n>t1\n" + "GGCA" * 150 + "\n")
    fq = tmp_path / "r.fq"; fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 10, "I" * 40) for i in range(8)))
    log = tmp_path / "log.txt"
    env = dict(os.environ, SFB200_STUB_LOG=str(log))
    idx = tmp_path / "idx"
    subprocess.check_call([exe, "index", "-t", str(fa), "-o", str(idx), "-k", "31", "--saveDeviceIndex"], env=env, stderr=subprocess.DEVNULL)
    calls = [c.split()[0] for c in log.read_text().strip().split("\n")]
    assert calls == ["index_build", "index_save"] and (idx / "device_index.bin").exists() and (idx / "versionInfo.json").exists()
    log.unlink()
    subprocess.check_call([exe, "quant", "-i", str(idx), "-l", "U", "-r", str(fq), "-o", str(tmp_path / "o")], env=env, stderr=subprocess.DEVNULL)
    calls = [c.split()[0] for c in log.read_text().strip().split("\n")]
    assert calls[0] == "index_load" and "index_build" not in calls and (tmp_path / "o" / "quant.sf").exists()
    # without the file the index is rebuilt from the stored sequence
    (idx / "device_index.bin").unlink(); log.unlink()
    subprocess.check_call([exe, "quant", "-i", str(idx), "-l", "U", "-r", str(fq), "-o", str(tmp_path / "o2")], env=env, stderr=subprocess.DEVNULL)
    assert log.read_text().split()[0] == "index_build"


def test_driver_host_flow_with_stub_device(tmp_path):
    """sfb200_quant.cpp linked against tests/stub_sfb200.cpp (a test double of the C ABI: canned device results, call log): the
    driver's HOST flow -- effective lengths, FLD hand-over to the bias model, corrected lengths in quant.sf, every aux file --
    without a GPU.  The real thing runs in the gpu tests below."""
    exe = str(tmp_path / "sfb200-quant-stub")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "sailfish_b200", "host", "sfb200_quant.cpp"),
                           os.path.join(ROOT, "tests", "stub_sfb200.cpp"), "-lz"])
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 100 + "\n>t1\n" + "GGCA" * 150 + "\n>t2\n" + "TTGA" * 60 + "\n")
    for tag in "12":
        (tmp_path / ("r%s.fq" % tag)).write_text("".join("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 10, "I" * 40) for i in range(4)))
    base = [exe, "quant", "-t", str(fa), "-l", "IU", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq")]
    log = tmp_path / "log.txt"
    env = dict(os.environ, SFB200_STUB_LOG=str(log))

    def run(args, out):
        if log.exists():
            log.unlink()
        subprocess.check_call(base + ["-o", str(tmp_path / out)] + args, env=env, stderr=subprocess.DEVNULL)
        rows = [l.split("\t") for l in open(tmp_path / out / "quant.sf").read().strip().split("\n")[1:]]
        return log.read_text().strip().split("\n"), rows, json.load(open(tmp_path / out / "aux" / "meta_info.json"))

    def vec(out, name, dt):
        return np.frombuffer(gzip.open(tmp_path / out / "aux" / name).read(), dtype=dt)

    # plain run: smoothed effective lengths from the observed FLD (12000 samples on 180..219), no bias calls, every aux file
    calls, rows, meta = run(["--dumpEq", "--numBootstraps", "2"], "o1")
    assert [c.split()[0] for c in calls] == ["index_build", "map_begin", "map_batch", "em_run"]
    assert calls[3] == "em_run 3 vb=0 eff0=201.500000 eff1=401.500000"
    assert [r[0] for r in rows] == ["t0", "t1", "t2"] and [r[2] for r in rows] == ["201.5", "401.5", "41.5"] and [r[4] for r in rows] == ["2", "1", "0"]
    assert meta["frag_dist_length"] == 999 and meta["bias_correct"] is False and meta["num_bias_bins"] == 4096 and meta["samp_type"] == "bootstrap"
    assert meta["num_processed"] == 4 and meta["num_mapped"] == 3 and meta["em_iterations"] == 51 and "start_time" in meta
    real = vec("o1", "fld.gz", np.int32)
    assert len(real) == 1000 and real.sum() == 10000 and real[:180].sum() == 0 and real[219:].sum() == 0      # bin 219 is cut by the 1 - 1e-6 rule
    assert (vec("o1", "observed_bias.gz", np.int32) == 1).all() and len(vec("o1", "observed_gc.gz", np.int32)) == 101
    assert (vec("o1", "expected_bias.gz", np.float64) == 1.0).all() and len(vec("o1", "expected_gc.gz", np.float64)) == 101
    boots = np.frombuffer(gzip.open(tmp_path / "o1" / "aux" / "bootstrap" / "bootstraps.gz").read(), dtype=np.float64)
    assert boots.tolist() == [1.0] * 3 + [2.0] * 3
    cmd = json.load(open(tmp_path / "o1" / "cmd_info.json"))                                                     # SailfishQuantify.cpp:1262-1276
    assert cmd["transcripts"] == str(fa) and cmd["libType"] == "IU" and cmd["mates1"] == str(tmp_path / "r1.fq") and cmd["output"] == str(tmp_path / "o1")
    assert cmd["dumpEq"] == [] and cmd["numBootstraps"] == "2" and list(cmd)[0] == "sf_version"
    # --unsmoothedFLD
    calls, rows, meta = run(["--unsmoothedFLD", "--useVBOpt"], "o2")
    assert calls[3].startswith("em_run 3 vb=1 eff0=202.0000")                                                  # mean of 180..218, float pdf
    # --gcBiasCorrect: samples requested before the first batch; the model carries the FLD table, the strand tallies, both observed arrays
    calls, rows, meta = run(["--gcBiasCorrect", "--gcSpeedSamp", "3", "--numBiasSamples", "1234"], "o3")
    assert [c.split()[0] for c in calls] == ["index_build", "map_begin", "map_set_bias", "map_batch", "em_run_bias"]
    assert calls[2] == "map_set_bias 0 1 1234"
    assert calls[4].startswith("em_run_bias 3 mode=2 gc_samp=3 fwd=2 rc=1 n_cdf=219 fld_max=999 rb7=8 og100=101 cdf_last=1.0000")
    assert [r[2] for r in rows] == ["100.75", "200.75", "20.75"]                                               # the corrected lengths (stub: half)
    assert meta["bias_correct"] is False                                                                       # opts.biasCorrect only (GZipWriter.cpp:178)
    assert vec("o3", "observed_gc.gz", np.int32).tolist() == list(range(1, 102)) and vec("o3", "observed_bias.gz", np.int32)[:3].tolist() == [1, 2, 3]
    # --biasCorrect with too few sampled fragment lengths (--numFragSamples above what was seen): the prior normal is the FLD
    calls, rows, meta = run(["--biasCorrect", "--numFragSamples", "20000"], "o4")
    assert calls[2] == "map_set_bias 1 0 1000000" and " mode=1 gc_samp=1 " in calls[4] and meta["bias_correct"] is True
    n_cdf = int(calls[4].split("n_cdf=")[1].split()[0])
    assert 400 < n_cdf < 700                                                                                   # N(200, 80) truncated at 1 - 1e-6
    real = vec("o4", "fld.gz", np.int32)
    assert real.sum() == 10000 and abs(float((real * np.arange(1000)).sum()) / 10000 - 200) < 5


def test_device_parse_feeding_with_stub_device(tmp_path):
    """--deviceParse: the driver reads raw FASTQ text block-wise and carries what sfb200_map_fastq did not consume over to the next
    block.  With the stub (tests/stub_sfb200.cpp: counts complete four-line records and checks that every block starts with '@'):
    several files per mate, a last line without its newline, CRLF, tiny blocks, mates whose records differ in size."""
    exe = str(tmp_path / "sfb200-quant-stub")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "sailfish_b200", "host", "sfb200_quant.cpp"),
                           os.path.join(ROOT, "tests", "stub_sfb200.cpp"), "-lz"])
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 100 + "\n")
    rng = np.random.default_rng(8)

    def fastq(path, n, first, eol="\n", final_newline=True, long_names=False):
        recs = []
        for i in range(n):
            L = int(rng.integers(30, 120))
            name = "@read%d%s" % (first + i, " a much longer description of this read" * 3 if long_names else "")
            recs.append(name + eol + "ACGT" * (L // 4) + eol + "+" + eol + "@" * (L // 4 * 4) + eol)
        text = "".join(recs)
        path.write_text(text if final_newline else text[:-len(eol)], newline="")

    fastq(tmp_path / "a1.fq", 700, 0); fastq(tmp_path / "a2.fq", 700, 0, long_names=True)              # mate 2 records are 4x larger
    fastq(tmp_path / "b1.fq", 300, 700, final_newline=False); fastq(tmp_path / "b2.fq", 300, 700, eol="\r\n")
    log = tmp_path / "log.txt"
    env = dict(os.environ, SFB200_STUB_LOG=str(log))
    for block in ("0", "4096", "700", "100"):                          # 100: every record is longer than a block (the carried text outgrows the room in front)
        if log.exists():
            log.unlink()
        subprocess.check_call([exe, "quant", "-t", str(fa), "-l", "IU", "-1", str(tmp_path / "a1.fq"), str(tmp_path / "b1.fq"), "-2", str(tmp_path / "a2.fq"),
                               str(tmp_path / "b2.fq"), "-o", str(tmp_path / ("o" + block)), "--deviceParse", "--blockBytes", block], env=env, stderr=subprocess.DEVNULL)
        calls = [l for l in log.read_text().strip().split("\n") if l.startswith("map_fastq")]
        assert not any("BAD_START" in c for c in calls)
        assert sum(int(c.split()[1]) for c in calls) == 1000 and all(c.endswith("paired=1") for c in calls)
        assert sum(int(c.split()[2]) for c in calls) == (tmp_path / "a1.fq").stat().st_size + (tmp_path / "b1.fq").stat().st_size + 1     # + the appended newline
        assert json.load(open(tmp_path / ("o" + block) / "aux" / "meta_info.json"))["num_processed"] == 1000
        if block != "0":
            assert len(calls) > 10
    # blank lines at the end of a file (some writers leave them; the host parser skips them) must not shift the next file's records
    (tmp_path / "c1.fq").write_text((tmp_path / "a1.fq").read_text() + "\n\n", newline="")
    (tmp_path / "c2.fq").write_text((tmp_path / "b2.fq").read_bytes().decode() + "\r\n", newline="")
    log.unlink()
    subprocess.check_call([exe, "quant", "-t", str(fa), "-l", "U", "-r", str(tmp_path / "c1.fq"), str(tmp_path / "c2.fq"), str(tmp_path / "a1.fq"),
                           "-o", str(tmp_path / "ob"), "--deviceParse", "--blockBytes", "0"], env=env, stderr=subprocess.DEVNULL)
    calls = [l for l in log.read_text().strip().split("\n") if l.startswith("map_fastq")]
    assert not any("BAD_START" in c for c in calls) and sum(int(c.split()[1]) for c in calls) == 700 + 300 + 700
    # single-end; a truncated file and mates of different length are errors
    log.unlink()
    subprocess.check_call([exe, "quant", "-t", str(fa), "-l", "U", "-r", str(tmp_path / "a1.fq"), "-o", str(tmp_path / "os"), "--deviceParse", "--blockBytes", "5000"],
                          env=env, stderr=subprocess.DEVNULL)
    assert sum(int(l.split()[1]) for l in log.read_text().strip().split("\n") if l.startswith("map_fastq")) == 700
    whole = (tmp_path / "a1.fq").read_text()
    (tmp_path / "trunc.fq").write_text(whole[:whole.rindex("@read") + 20])                      # the last record stops inside its sequence line
    r = subprocess.run([exe, "quant", "-t", str(fa), "-l", "U", "-r", str(tmp_path / "trunc.fq"), "-o", str(tmp_path / "ot"), "--deviceParse"], env=env, capture_output=True)
    assert r.returncode == 1 and b"truncated record" in r.stderr
    r = subprocess.run([exe, "quant", "-t", str(fa), "-l", "IU", "-1", str(tmp_path / "a1.fq"), "-2", str(tmp_path / "b2.fq"), "-o", str(tmp_path / "om"), "--deviceParse"],
                       env=env, capture_output=True)
    assert r.returncode == 1 and b"different numbers of reads" in r.stderr
    with gzip.open(tmp_path / "z.fq.gz", "wt") as f:
        f.write((tmp_path / "a1.fq").read_text())
    r = subprocess.run([exe, "quant", "-t", str(fa), "-l", "U", "-r", str(tmp_path / "z.fq.gz"), "-o", str(tmp_path / "oz"), "--deviceParse"], env=env, capture_output=True)
    assert r.returncode == 1 and b"inflate gzipped files first" in r.stderr


def test_bias_option_checks(tmp_path):
    """SailfishQuantify.cpp:1293-1309: the two corrections exclude each other; GC correction is switched off for single-end libraries"""
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 30 + "\n")
    fq = tmp_path / "r.fq"; fq.write_text("@r\n" + "ACGT" * 10 + "\n+\n" + "I" * 40 + "\n")
    base = [build_exe(), "quant", "-t", str(fa), "-l", "U", "-r", str(fq), "-o", str(tmp_path / "o")]
    r = subprocess.run(base + ["--biasCorrect", "--gcBiasCorrect"], capture_output=True)
    assert r.returncode == 1 and b"simultaneously is not yet supported" in r.stderr
    r = subprocess.run(base + ["--gcBiasCorrect"], capture_output=True)
    assert b"only implemented for paired-end libraries" in r.stderr
    r = subprocess.run(base + ["--biasCorrect", "--noEffectiveLengthCorrection"], capture_output=True)
    assert r.returncode == 2 or b"need the effective length correction" in r.stderr + r.stdout


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = tmp_path / "t.fa"; fa.write_text(">t0\n" + "ACGT" * 30 + "\n")
    fq = tmp_path / "r.fq"; fq.write_text("@r\n" + "ACGT" * 10 + "\n+\n" + "I" * 40 + "\n")
    r = subprocess.run([build_exe(), "-t", str(fa), "-l", "U", "-r", str(fq), "-o", str(tmp_path / "o")], capture_output=True)
    assert r.returncode == 3 and b"no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_sample_data_end_to_end_cpp(sample_data, tmp_path):
    d = sample_data
    seqs = split_seqs(d["txp_seq"], d["txp_len"])
    names = [str(n) for n in d["names"]]
    fa = tmp_path / "transcripts.fasta"
    with open(fa, "w") as f:
        for n, s in zip(names, seqs):
            s = s.decode()
            f.write(">%s extra words\n%s\n" % (n, "\n".join(s[j:j + 60] for j in range(0, len(s), 60))))
    for tag, reads, off in (("1", d["reads1"], d["off1"]), ("2", d["reads2"], d["off2"])):
        with gzip.open(tmp_path / ("reads_%s.fastq.gz" % tag), "wt") as f:
            for i in range(len(off) - 1):
                s = reads[int(off[i]):int(off[i + 1])].tobytes().decode()
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))
    out = tmp_path / "q"
    subprocess.check_call([build_exe(), "quant", "-t", str(fa), "-l", "IU", "-1", str(tmp_path / "reads_1.fastq.gz"),
                           "-2", str(tmp_path / "reads_2.fastq.gz"), "-o", str(out), "--dumpEq", "--numBootstraps", "3", "--batchReads", "3000"])
    lines = open(out / "quant.sf").read().strip().split("\n")
    assert lines[0] == "Name\tLength\tEffectiveLength\tTPM\tNumReads" and len(lines) == 1 + len(names)
    rows = [l.split("\t") for l in lines[1:]]
    assert [r[0] for r in rows] == names and [int(r[1]) for r in rows] == [int(x) for x in d["txp_len"]]
    np.testing.assert_allclose([float(r[4]) for r in rows], d["ref_est_vb0"], rtol=1.2e-4, atol=1e-6)   # 1e-4 parity + %g's 6 digits
    np.testing.assert_allclose([float(r[2]) for r in rows], d["eff"], rtol=1e-5)
    assert abs(sum(float(r[3]) for r in rows) - 1e6) < 50
    meta = json.load(open(out / "aux" / "meta_info.json"))
    assert meta["num_processed"] == 10000 and meta["num_mapped"] == int(d["num_mapped"]) and meta["samp_type"] == "bootstrap"
    eq = open(out / "aux" / "eq_classes.txt").read().split("\n")
    assert int(eq[0]) == len(names) and int(eq[1]) == len(d["counts"])
    boots = np.frombuffer(gzip.open(out / "aux" / "bootstrap" / "bootstraps.gz").read(), dtype=np.float64).reshape(3, len(names))
    np.testing.assert_allclose(boots.sum(axis=1), float(d["num_mapped"]), rtol=1e-9)
    # the Gibbs path and VBEM through the same command, from an index directory
    out2 = tmp_path / "q2"
    subprocess.check_call([EXE, "index", "-t", str(fa), "-o", str(tmp_path / "idx"), "-k", "31"])
    (tmp_path / "t2g.tsv").write_text("".join("%s\tgene%d\n" % (n, i // 5) for i, n in enumerate(names)))
    subprocess.check_call([EXE, "quant", "-i", str(tmp_path / "idx"), "-l", "IU", "-1", str(tmp_path / "reads_1.fastq.gz"),
                           "-2", str(tmp_path / "reads_2.fastq.gz"), "-o", str(out2), "--useVBOpt", "--numGibbsSamples", "4", "-g", str(tmp_path / "t2g.tsv")])
    rows2 = [l.split("\t") for l in open(out2 / "quant.sf").read().strip().split("\n")[1:]]
    np.testing.assert_allclose([float(r[4]) for r in rows2], d["ref_est_vb1"], rtol=1.2e-4, atol=1e-6)
    genes = [l.split("\t") for l in open(out2 / "quant.genes.sf").read().strip().split("\n")[1:]]
    assert [g[0] for g in genes] == ["gene%d" % i for i in range((len(names) + 4) // 5)]
    assert abs(sum(float(g[4]) for g in genes) - sum(float(r[4]) for r in rows2)) < 1e-3 * float(d["num_mapped"])
    gib = np.frombuffer(gzip.open(out2 / "aux" / "bootstrap" / "bootstraps.gz").read(), dtype=np.int32).reshape(4, len(names))
    assert (gib.sum(axis=1) == int(d["num_mapped"])).all()




@pytest.mark.gpu
@pytest.mark.parametrize("flag,mode", [("--biasCorrect", 1), ("--gcBiasCorrect", 2)])
def test_sample_data_bias_correction_cpp(sample_data, tmp_path, flag, mode):
    """`sfb200-quant quant --biasCorrect / --gcBiasCorrect` against the oracle's pipeline: samples collected while mapping, effective
    lengths recomputed inside the optimizer, corrected lengths in quant.sf"""
    from oracle import pyoracle as O
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
    out = tmp_path / "q"
    subprocess.check_call([build_exe(), "quant", "-t", str(fa), "-l", "IU", "-1", str(tmp_path / "reads_1.fastq"), "-2", str(tmp_path / "reads_2.fastq"),
                           "-o", str(out), flag, "--fldMean", "200", "--fldSD", "80", "--batchReads", "3000"])
    rows = [l.split("\t") for l in open(out / "quant.sf").read().strip().split("\n")[1:]]
    # the oracle's run of the same pipeline
    oix = O.Index(seqs, k=31)
    fmt = O.parse_libtype("IU")
    run = O.Run(oix, O.MapOpts.default(fmt))
    run.set_bias(mode == 1, mode == 2, 1000000)
    run.map_batch(d["reads1"].tobytes(), d["off1"], d["reads2"].tobytes(), d["off2"], n_threads=4)
    w = run.finish()
    rb, og = run.finish_bias()
    assert int(w["fld"].sum()) < 10000                               # too few sampled fragments: the prior normal is the FLD (:966-976)
    x = np.arange(1000, dtype=np.float64)
    dens = np.exp(-0.5 * ((x - 200.0) / 80.0) ** 2) / 80.0
    fld = np.floor(dens * 10000 / dens.sum() + 0.5).astype(np.uint32)
    rc, want, eff_want, it_o, _ = O.em_run_bias(mode, seqs, w["row_ptr"], w["labels"], w["counts"], d["eff"], int(w["counters"][1]),
                                                int(w["counters"][4]), int(w["counters"][5]), rb, og, fld)
    assert rc == 0 and it_o >= 50
    np.testing.assert_allclose([float(r[4]) for r in rows], want, rtol=1.2e-4, atol=1e-6)
    np.testing.assert_allclose([float(r[2]) for r in rows], eff_want, rtol=1e-5)
    assert (np.abs(eff_want - np.maximum(d["eff"], 1.0)) > 1e-6).sum() > 10
    assert json.load(open(out / "aux" / "meta_info.json"))["bias_correct"] is (flag == "--biasCorrect")     # opts.biasCorrect (GZipWriter.cpp:178)
    og_file = np.frombuffer(gzip.open(out / "aux" / "observed_gc.gz").read(), dtype=np.int32)
    rb_file = np.frombuffer(gzip.open(out / "aux" / "observed_bias.gz").read(), dtype=np.int32)
    assert og_file.tolist() == og.tolist() and rb_file.tolist() == rb.tolist()
