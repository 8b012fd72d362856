"""Property test of the FASTQ / FASTA ingestion of sfb200-quant (sailfish_b200/host/fastx_reader.hpp) against a plain Python
parse: random records, line endings, final newline, block sizes, batch sizes and thread counts (CPU only, --parseOnly)."""
import json
import os
import subprocess

from hypothesis import given, settings, strategies as st

from test_host_quant_cli import build_exe, fnv

seq = st.text(alphabet="ACGTNacgt", min_size=0, max_size=70)


@settings(max_examples=40, deadline=None)
@given(seqs=st.lists(seq, min_size=1, max_size=60), crlf=st.booleans(), final_nl=st.booleans(), fasta=st.booleans(),
       block=st.sampled_from([16, 17, 33, 100, 1000, 0]), batch=st.integers(1, 70), threads=st.integers(1, 4), trailing_blank=st.booleans())
def test_parser_matches_python(tmp_path_factory, seqs, crlf, final_nl, fasta, block, batch, threads, trailing_blank):
    d = tmp_path_factory.mktemp("fz")
    eol = "\r\n" if crlf else "\n"
    if not final_nl and not seqs[-1]:
        seqs = seqs[:-1] + ["A"]                              # an empty last read without a final newline is a three-line record: rejected as truncated
    if fasta:
        seqs = [s or "A" for s in seqs]                       # an empty FASTA record is a header followed by another header: keep it simple
        recs = [">r%d x" % i + eol + eol.join(s[j:j + 13] for j in range(0, len(s), 13)) for i, s in enumerate(seqs)]
    else:
        recs = ["@r%d x" % i + eol + s + eol + "+" + eol + "I" * len(s) for i, s in enumerate(seqs)]
    txt = eol.join(recs) + (eol if final_nl else "") + (eol + eol if trailing_blank and final_nl else "")
    p = d / ("r.fa" if fasta else "r.fq")
    with open(p, "w", newline="") as f:
        f.write(txt)
    out = subprocess.check_output([build_exe(), "--parseOnly", "-r", str(p), "--blockBytes", str(block), "--batchReads", str(batch), "-p", str(threads)])
    got = json.loads(out)
    assert got["records"] == len(seqs) and got["bases1"] == sum(map(len, seqs))
    assert got["fnv1a"] == fnv(seqs, [])


def test_device_fastq_extraction_arithmetic(tmp_path):
    """sailfish_b200/csrc/fastq_core.inl (the per-chunk bodies of k_fq_count / k_fq_mark / k_fq_copy) compiled as host code and
    replayed over random FASTQ text -- see tests/fastq_core_test.cpp"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fastq_core_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(root, "tests", "fastq_core_test.cpp")])
    assert "fastq core ok" in subprocess.check_output([exe]).decode()
