"""`sailfish quant`-shaped driver on top of the C ABI: transcripts + reads -> quant.sf (and aux/).

This is the plumbing around the hot path, kept deliberately thin: FASTA/FASTQ parsing in Python,
everything between "a batch of reads" and "per-transcript counts" on the GPU through libsfb200.  Output formats follow the
reference writers:

  quant.sf              Name, Length, EffectiveLength, TPM, NumReads; doubles printed like cppformat's `{}` == printf("%g")
                        (reference src/GZipWriter.cpp:194-248)
  aux/eq_classes.txt    T, E, T names, then per class `n \\t tid_1 .. tid_n \\t count` (src/GZipWriter.cpp:51-92), with --dumpEq
  aux/bootstrap/bootstraps.gz   raw little-endian f64 rows (bootstraps) or i32 rows (Gibbs) (src/GZipWriter.cpp:250-284)
  aux/meta_info.json    the counters of src/GZipWriter.cpp:163-190 that exist on this path

    python -m sailfish_b200.quant -t transcripts.fasta -l IU -1 reads_1.fastq -2 reads_2.fastq -o out_dir
"""
import argparse
import gzip
import json
import os
import sys
import time

import numpy as np

from . import capi, efflen

# LibraryFormat ids (reference include/LibraryFormat.hpp:7-9,89-98): type | orientation << 1 | strandedness << 3
_SE, _PE = 0, 1
_SAME, _AWAY, _TOWARD, _NONE = 0, 1, 2, 3
_SA, _AS, _S, _A, _U = 0, 1, 2, 3, 4


def parse_library_format(s):
    """parseLibraryFormatString of the CLI (reference src/SailfishUtils.cpp:63-97) -> LibraryFormat::formatID()"""
    table = {"IU": (_PE, _TOWARD, _U), "ISF": (_PE, _TOWARD, _SA), "ISR": (_PE, _TOWARD, _AS),
             "OU": (_PE, _AWAY, _U), "OSF": (_PE, _AWAY, _SA), "OSR": (_PE, _AWAY, _AS),
             "MU": (_PE, _SAME, _U), "MSF": (_PE, _SAME, _S), "MSR": (_PE, _SAME, _A),
             "U": (_SE, _NONE, _U), "SF": (_SE, _NONE, _S), "SR": (_SE, _NONE, _A)}
    key = s.upper()
    if key not in table:
        raise ValueError("unknown library type %r" % s)
    t, o, st = table[key]
    return (t & 1) | ((o & 3) << 1) | ((st & 7) << 3)


def _open(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "rt")


def read_fasta(path):
    names, seqs, cur = [], [], []
    with _open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if cur or names:
                    seqs.append("".join(cur)); cur = []
                names.append(line[1:].split()[0])
            elif line:
                cur.append(line)
    if names:
        seqs.append("".join(cur))
    return names, seqs


def read_fastx_batches(path, batch):
    """yields lists of sequences (FASTQ or FASTA reads), `batch` at a time"""
    out = []
    with _open(path) as f:
        first = f.readline()
        if not first:
            return
        fastq = first.startswith("@")
        if fastq:
            while first:
                out.append(f.readline().strip())
                f.readline(); f.readline()
                if len(out) == batch:
                    yield out; out = []
                first = f.readline()
        else:
            cur = []
            for line in f:
                line = line.strip()
                if line.startswith(">"):
                    out.append("".join(cur)); cur = []
                    if len(out) == batch:
                        yield out; out = []
                elif line:
                    cur.append(line)
            out.append("".join(cur))
    if out:
        yield out


def fmt_g(x):
    return "%g" % x


def write_quant_sf(path, names, lengths, eff, alphas, num_mapped):
    """GZipWriter::writeAbundances (reference src/GZipWriter.cpp:194-248)"""
    tpm = capi.tpm(alphas, eff, num_mapped) if num_mapped > 0 and np.sum(alphas) > 0 else np.zeros(len(alphas))
    with open(path, "w") as f:
        f.write("Name\tLength\tEffectiveLength\tTPM\tNumReads\n")
        for i, n in enumerate(names):
            f.write("%s\t%d\t%s\t%s\t%s\n" % (n, lengths[i], fmt_g(eff[i]), fmt_g(tpm[i]), fmt_g(alphas[i])))
    return tpm


def write_eq_classes(path, names, row_ptr, labels, counts):
    """GZipWriter::writeEquivCounts (reference src/GZipWriter.cpp:51-92)"""
    with open(path, "w") as f:
        f.write("%d\n%d\n" % (len(names), len(counts)))
        for n in names:
            f.write(n + "\n")
        for e in range(len(counts)):
            ids = labels[int(row_ptr[e]):int(row_ptr[e + 1])]
            f.write("%d\t%s\t%d\n" % (len(ids), "\t".join(str(int(t)) for t in ids), int(counts[e])))


def map_fastq_files(ctx, files1, files2=None, block=64 << 20):
    """--deviceParse: FASTQ text goes to the GPU as it is (capi.Context.map_fastq); what a call did not consume -- an incomplete
    record, or records the other mate had no partner for yet -- is put in front of the next block.  Plain four-line FASTQ only."""
    def stream(files):
        for path in files:
            with open(path, "rb") as f:
                if f.read(2) == b"\x1f\x8b":
                    raise ValueError("%s: --deviceParse reads plain FASTQ text; inflate gzipped files first (or drop the option)" % path)
                f.seek(0)
                last = b"\n"
                while True:
                    chunk = f.read(block)
                    if not chunk:
                        break
                    last = chunk[-1:]
                    yield chunk
                if last != b"\n":
                    yield b"\n"                                   # a file whose last line lacks its newline gets one
    paired = files2 is not None
    s1, s2 = stream(files1), stream(files2) if paired else iter(())
    buf1 = buf2 = b""
    done1 = done2 = False
    want, total = block, 0
    while True:
        while len(buf1) < want and not done1:
            nxt = next(s1, None)
            done1 = nxt is None
            buf1 += nxt or b""
        while paired and len(buf2) < want and not done2:
            nxt = next(s2, None)
            done2 = nxt is None
            buf2 += nxt or b""
        if not buf1 and not buf2:
            break
        n = c1 = c2 = 0
        if buf1 and (not paired or buf2):
            n, c1, c2 = ctx.map_fastq(buf1, buf2 if paired else None)
        if n == 0:
            if done1 and (not paired or done2):
                if buf1.strip() or (paired and buf2.strip()):
                    raise ValueError("mate files hold different numbers of reads" if paired and bool(buf1.strip()) != bool(buf2.strip())
                                     else "truncated record at end of file")
                break
            if paired:
                out1, out2 = done1 and not buf1.strip(), done2 and not buf2.strip()      # a mate with nothing left at all
                if out1 != out2 and (buf2.strip() if out1 else buf1.strip()):
                    raise ValueError("mate files hold different numbers of reads")
            want += block
            continue
        want = block
        total += n
        buf1 = buf1[c1:]
        if paired:
            buf2 = buf2[c2:]
    return total


def write_aux_vectors(aux, fld_counts, obs_bias, obs_gc):
    """the binary vectors of GZipWriter::writeMeta (src/GZipWriter.cpp:139-161): raw little-endian elements, gzip.  fld.gz is a
    realisation of 10000 draws from the fragment length pdf (random in the reference, fixed seed here; skipped when there is no
    distribution); the two "expected" vectors are never filled by the reference and stay all ones"""
    if fld_counts is not None:
        pdf = np.zeros(len(fld_counts), np.float64)
        cdf, _ = efflen.empirical_cdf(fld_counts)
        pdf[:len(cdf)] = np.diff(np.concatenate([[0.0], cdf.astype(np.float64)]))
        pdf = np.where(np.isfinite(pdf) & (pdf > 0), pdf, 0.0)
        samples = np.zeros(len(pdf), np.int32)
        if pdf.sum() > 0:
            draws = np.random.default_rng(0x5f3759df).choice(len(pdf), size=10000, p=pdf / pdf.sum())
            samples = np.bincount(draws, minlength=len(pdf)).astype(np.int32)
        with gzip.open(os.path.join(aux, "fld.gz"), "wb") as f:
            f.write(samples.tobytes())
    for fn, arr in (("expected_bias.gz", np.ones(4096, np.float64)), ("observed_bias.gz", np.asarray(obs_bias).astype(np.int32)),
                    ("expected_gc.gz", np.ones(101, np.float64)), ("observed_gc.gz", np.asarray(obs_gc).astype(np.int32))):
        with gzip.open(os.path.join(aux, fn), "wb") as f:
            f.write(np.ascontiguousarray(arr).tobytes())


def quantify(transcripts, reads1, reads2=None, libtype=None, out_dir="sailfish_quant", k=31, use_vb=False, n_boot=0, n_gibbs=0,
             dump_eq=False, batch=1_000_000, device=0, no_eff_len_correction=False, map_kw=None, bias_correct=False,
             gc_bias_correct=False, num_bias_samples=1000000, gc_speed_samp=1, unsmoothed_fld=False, device_parse=False,
             block_bytes=64 << 20):
    t_start = time.time()
    names, seqs = read_fasta(transcripts)
    lengths = np.array([len(s) for s in seqs], np.uint32)
    paired = reads2 is not None
    fmt = parse_library_format(libtype or ("IU" if paired else "U"))
    if paired != bool(fmt & 1):
        raise ValueError("library type %s does not match the number of read files" % libtype)
    capi.bind_host_near_device(device)          # reader buffers on the GPU's NUMA node (one process per GPU)
    ctx = capi.Context(device)
    ctx.index_build(seqs=seqs, k=k)
    ctx.map_begin(capi.MapOpts.default(fmt, **(map_kw or {})))
    if bias_correct and gc_bias_correct:                                      # SailfishQuantify.cpp:1293-1297
        raise ValueError("Enabling both sequence-specific and fragment GC bias correction simultaneously is not yet supported.")
    if gc_bias_correct and not paired:                                        # :1298-1309
        print("Fragment GC bias correction is currently only implemented for paired-end libraries. It is being disabled", file=sys.stderr)
        gc_bias_correct = False
    do_bias = bias_correct or gc_bias_correct
    if do_bias and no_eff_len_correction:
        raise ValueError("bias correction needs the effective length correction")
    if do_bias:
        ctx.map_set_bias(bias_correct, gc_bias_correct, num_bias_samples)
    it2 = read_fastx_batches(reads2, batch) if paired and not device_parse else None
    if device_parse:
        map_fastq_files(ctx, [reads1], [reads2] if paired else None, block_bytes)
    for r1 in (() if device_parse else read_fastx_batches(reads1, batch)):
        b1, o1 = capi.pack_reads(r1)
        if paired:
            r2 = next(it2)
            if len(r2) != len(r1):
                raise ValueError("mate files have different numbers of reads")
            b2, o2 = capi.pack_reads(r2)
            ctx.map_batch(b1, o1, b2, o2)
        else:
            ctx.map_batch(b1, o1)
    g = ctx.map_finish()
    n_clip = ctx.map_clipped()
    if n_clip:
        print("WARNING: %d mates are longer than 256 bases and were mapped by their first 256 (the reference maps the whole read)" % n_clip,
              file=sys.stderr)
    counters = g["counters"]
    num_mapped = int(counters[1])
    eff = efflen.effective_lengths(lengths, g["fld"], max_frag_len=ctx.map_opts.max_frag_len,
                                   num_frag_samples=ctx.map_opts.num_frag_samples, single_end=not paired,
                                   no_correction=no_eff_len_correction, unsmoothed=unsmoothed_fld)
    os.makedirs(os.path.join(out_dir, "aux"), exist_ok=True)
    if dump_eq:
        rp, lab, cnt = ctx.eq_export()
        write_eq_classes(os.path.join(out_dir, "aux", "eq_classes.txt"), names, rp, lab, cnt)
    if do_bias:
        # readExp.setFragLengthDist (:966-984, :1039) -> EmpiricalDistribution's cdf table; the optimizer recomputes the effective
        # lengths at iterations 50 / 500 / 1000 and quant.sf reports the corrected ones (CollapsedEMOptimizer.cpp:820-840, :888)
        enough = paired and int(np.asarray(g["fld"], np.uint64).sum()) >= ctx.map_opts.num_frag_samples
        fld_counts = g["fld"] if enough else efflen.normal_frag_length_counts(ctx.map_opts.max_frag_len, ctx.map_opts.num_frag_samples)
        cdf, fld_max = efflen.empirical_cdf(fld_counts)
        rb, og = ctx.map_get_bias()
        alphas, eff, iters, mrd = ctx.em_run_bias(2 if gc_bias_correct else 1, eff, num_mapped, int(counters[4]), int(counters[5]), rb, og,
                                                  cdf, fld_max, gc_samp=gc_speed_samp, opts=capi.EMOpts.default(use_vb=int(use_vb)))
    else:
        alphas, iters, mrd = ctx.em_run(eff, num_mapped, capi.EMOpts.default(use_vb=int(use_vb)))
    write_quant_sf(os.path.join(out_dir, "quant.sf"), names, lengths, eff, alphas, num_mapped)
    counts = None
    if not no_eff_len_correction:
        enough = paired and int(np.asarray(g["fld"], np.uint64).sum()) >= ctx.map_opts.num_frag_samples
        counts = g["fld"] if enough else efflen.normal_frag_length_counts(ctx.map_opts.max_frag_len, ctx.map_opts.num_frag_samples)
    obs_bias, obs_gc = (ctx.map_get_bias() if do_bias else (np.ones(4096, np.uint32), np.ones(101, np.uint32)))
    write_aux_vectors(os.path.join(out_dir, "aux"), counts, obs_bias, obs_gc)
    samp_type = "none"
    if n_boot or n_gibbs:
        os.makedirs(os.path.join(out_dir, "aux", "bootstrap"), exist_ok=True)
        with gzip.open(os.path.join(out_dir, "aux", "bootstrap", "names.tsv.gz"), "wt") as f:
            f.write("\t".join(names) + "\n")
        if n_boot:
            rows = ctx.bootstrap_run(eff, n_boot, opts=capi.EMOpts.default(use_vb=int(use_vb)))
            samp_type = "bootstrap"
        else:
            rows = ctx.gibbs_run(eff, alphas / max(alphas.sum(), 1e-300), num_mapped, n_gibbs)
            samp_type = "gibbs"
        with gzip.open(os.path.join(out_dir, "aux", "bootstrap", "bootstraps.gz"), "wb") as f:
            f.write(np.ascontiguousarray(rows).tobytes())
    meta = {"sf_version": "0.10.0-b200", "samp_type": samp_type, "frag_dist_length": int(ctx.map_opts.max_frag_len) - 1,
            "bias_correct": bool(bias_correct), "num_bias_bins": 4096, "num_targets": len(names), "num_bootstraps": int(n_boot or n_gibbs),
            "num_processed": int(counters[0]), "num_mapped": num_mapped,
            "percent_mapped": 100.0 * num_mapped / max(int(counters[0]), 1), "call": "quant", "start_time": time.asctime(time.localtime(t_start)),
            "em_iterations": int(iters), "elapsed_s": time.time() - t_start}
    with open(os.path.join(out_dir, "aux", "meta_info.json"), "w") as f:
        json.dump(meta, f, indent=4)
    ctx.close()
    return dict(names=names, lengths=lengths, eff=eff, alphas=alphas, counters=counters, iters=iters, meta=meta)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="sailfish_b200.quant", description="transcript quantification on B200 (sailfish quant)")
    ap.add_argument("-t", "--transcripts", required=True, help="transcript FASTA (the index is built on the GPU from it)")
    ap.add_argument("-l", "--libType", default=None)
    ap.add_argument("-r", "--unmatedReads")
    ap.add_argument("-1", "--mates1")
    ap.add_argument("-2", "--mates2")
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("-k", "--kmerLen", type=int, default=31)
    ap.add_argument("--useVBOpt", action="store_true")
    ap.add_argument("--numBootstraps", type=int, default=0)
    ap.add_argument("--numGibbsSamples", type=int, default=0)
    ap.add_argument("--dumpEq", action="store_true")
    ap.add_argument("--noEffectiveLengthCorrection", action="store_true")
    ap.add_argument("--unsmoothedFLD", action="store_true")
    ap.add_argument("--deviceParse", action="store_true", help="FASTQ text parsed on the GPU (plain four-line FASTQ)")
    ap.add_argument("--blockBytes", type=int, default=64 << 20)
    ap.add_argument("--biasCorrect", action="store_true")
    ap.add_argument("--gcBiasCorrect", action="store_true")
    ap.add_argument("--numBiasSamples", type=int, default=1000000)
    ap.add_argument("--gcSpeedSamp", type=int, default=1)
    a = ap.parse_args(argv)
    if a.numBootstraps and a.numGibbsSamples:
        sys.exit("--numBootstraps and --numGibbsSamples are mutually exclusive (SailfishQuantify.cpp:1281-1287)")
    r1 = a.mates1 or a.unmatedReads
    if not r1:
        sys.exit("no reads given")
    res = quantify(a.transcripts, r1, a.mates2, a.libType, a.output, a.kmerLen, a.useVBOpt, a.numBootstraps, a.numGibbsSamples,
                   a.dumpEq, no_eff_len_correction=a.noEffectiveLengthCorrection, bias_correct=a.biasCorrect,
                   gc_bias_correct=a.gcBiasCorrect, num_bias_samples=a.numBiasSamples, gc_speed_samp=a.gcSpeedSamp,
                   unsmoothed_fld=a.unsmoothedFLD, device_parse=a.deviceParse, block_bytes=a.blockBytes)
    print(json.dumps(res["meta"]))


if __name__ == "__main__":
    main()
