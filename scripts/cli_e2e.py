"""End-to-end timing of the C++ driver from FASTQ text: synthetic transcriptome + single-end reads written as FASTQ to a
scratch directory, then `sfb200-quant quant`; prints one JSON line (reads/s from FASTQ, phases from the driver's log).
    python scripts/cli_e2e.py [--genes 40000] [--reads 4000000] [--dir /dev/shm/sfb200_cli]"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sailfish_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=40000)
ap.add_argument("--reads", type=int, default=4_000_000)
ap.add_argument("--dir", default="/dev/shm/sfb200_cli")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--device-parse", action="store_true", help="pass --deviceParse: FASTQ text parsed on the GPU (sfb200_map_fastq)")
ap.add_argument("--reuse", action="store_true", help="keep the files of an earlier invocation with the same sizes")
a = ap.parse_args()
os.makedirs(a.dir, exist_ok=True)
fa = os.path.join(a.dir, "t.fa")
fq = os.path.join(a.dir, "r.fq")
L = 76
have = a.reuse and os.path.exists(fa) and os.path.exists(fq) and os.path.getsize(fq) == a.reads * (11 + L + 3 + L + 1)
if have:
    ln = np.zeros(a.genes * 5, np.uint32)
else:
    seq, off, ln = synth.make_transcriptome(a.genes, seed=42)
with open(fa, "wb") if not have else open(os.devnull, "wb") as f:
    for i in range(len(ln) if not have else 0):
        f.write(b">t%d\n" % i)
        f.write(seq[int(off[i]):int(off[i]) + int(ln[i])].tobytes())
        f.write(b"\n")
with open(fq, "wb") if not have else open(os.devnull, "wb") as f:
    done = 0 if not have else a.reads
    c = 0
    while done < a.reads:
        n = min(1_000_000, a.reads - done)
        b1, _, _, _, _ = synth.make_reads(seq, off, ln, n, L, seed=1234, expr_seed=1234, stream=c)
        rec = np.empty((n, 11 + L + 3 + L + 1), np.uint8)              # "@r%08d\n" seq "\n+\n" qual "\n"
        ids = np.char.zfill(np.arange(done, done + n).astype(str), 8)
        rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
        rec[:, 2:10] = np.frombuffer("".join(ids).encode(), np.uint8).reshape(n, 8)
        rec[:, 10] = 10
        rec[:, 11:11 + L] = b1.reshape(n, L)
        rec[:, 11 + L] = 10; rec[:, 12 + L] = ord("+"); rec[:, 13 + L] = 10
        rec[:, 14 + L:14 + 2 * L] = ord("I")
        rec[:, -1] = 10
        f.write(rec.tobytes())
        done += n; c += 1
exe = os.path.join(ROOT, "sailfish_b200", "bin", "sfb200-quant")
cmd = [exe, "quant", "-t", fa, "-l", "U", "-r", fq, "-o", os.path.join(a.dir, "out")]
if a.threads:
    cmd += ["-p", str(a.threads)]
if a.device_parse:
    cmd += ["--deviceParse"]
t0 = time.time()
r = subprocess.run(cmd, capture_output=True, text=True)
dt = time.time() - t0
log = r.stderr
m_idx = re.search(r"index built in ([0-9.]+) s", log)
m_map = re.search(r"equivalence classes, ([0-9.]+) s", log)
m_tot = re.search(r"\(([0-9.]+) s in total\)", log)
out = {"rc": r.returncode, "device_parse": bool(a.device_parse), "reads": a.reads, "transcripts": int(len(ln)), "fastq_bytes": os.path.getsize(fq), "wall_s": dt,
       "transcripts_and_index_s": float(m_idx.group(1)) if m_idx else None, "ingest_and_map_s": float(m_map.group(1)) if m_map else None,
       "total_s": float(m_tot.group(1)) if m_tot else None}
if m_map:
    out["reads_per_s_from_fastq"] = a.reads / float(m_map.group(1))
    out["fastq_GB_per_s"] = os.path.getsize(fq) / float(m_map.group(1)) / 1e9
print(json.dumps(out))
if r.returncode:
    print(log, file=sys.stderr)
