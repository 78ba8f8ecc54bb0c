"""CPU tests of the drop-in boundary: libhcnv.so loads and exports every symbol include/hcnv.h declares, the
host-only B3 entry points (config, number formatting, text writers) behave like the reference's, and the
product path fails loudly without a GPU (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hadoopcnv_b200 as H
from hadoopcnv_b200 import _lib as L
from oracle import literal as Lit
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        L.build()
    return L.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "hcnv.h")).read()
    declared = set(re.findall(r"\b(hcnv_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hcnv_version() >= 100
    assert lib.hcnv_strerror(-5).decode().startswith("input not sorted")


def test_struct_layouts_match_header(lib):
    sizes = (C.c_int32 * 7)()
    lib.hcnv_abi_sizes(sizes)
    assert list(sizes) == [H.api.BLOCK_DTYPE.itemsize, H.api.LONG_BLOCK_DTYPE.itemsize, C.sizeof(L.BinParams),
                           C.sizeof(L.ContigInput), C.sizeof(L.Bins), C.sizeof(L.SegmentProblem), C.sizeof(L.Config)]


def test_struct_layouts_of_the_pipeline_structs_match_header(lib):
    sizes = (C.c_int32 * 4)()
    lib.hcnv_abi_sizes2(sizes)
    assert list(sizes) == [C.sizeof(L.ResultsDownload), C.sizeof(L.GenomeOutput), C.sizeof(L.SegmentPart), C.sizeof(L.Bins)]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(H.HcnvError) as e:
        H.Context(0)
    assert e.value.status == L.HCNV_E_CUDA


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hadoopcnv_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                for needle in ("import oracle", "from oracle", "hcnv_oracle", "oracle/", "orc_"):
                    assert needle not in src, (fn, needle)


CONFIG_EXAMPLE = """# it is assumed that this file exists on HDFS already
BAM_FILE: /bamfiles/child_hg19_SVs_v2.sort.bam
VCF_FILE: VCF/cov40.vcf
RUN_DEPTH_CALLER: true
RUN_GLOBAL_SORT: false
INIT_VCF_LOOKUP: true
RUN_BINNER: true
RUN_CNV_CALLER: true
RUN_SR_PE_EXTRACTER: false
BAM_FILE_REDUCERS: 800
TEXT_FILE_REDUCERS: 100
YARN_CONTAINER_MIN_MB: 1024
YARN_CONTAINER_MAX_MB: 8192
HEAPSIZE_MIN_MB: 100
HEAPSIZE_MAX_MB: 2000
MAPPER_MB: 2048
REDUCER_MB: 2048
REDUCER_TASKS: 24

# Configuration
ABBERATION_PENALTY: 0.01f
TRANSITION_PENALTY: 1.6f
"""


def test_config_load_example(tmp_path, lib):
    """Same keys/values as the reference's example/config.txt (':' separators, comments, the trailing 'f')."""
    p = tmp_path / "config.txt"
    p.write_text(CONFIG_EXAMPLE)
    cfg = H.config_load(str(p))
    assert cfg["BAM_FILE"] == "/bamfiles/child_hg19_SVs_v2.sort.bam" and cfg["VCF_FILE"] == "VCF/cov40.vcf"
    assert cfg["RUN_DEPTH_CALLER"] and not cfg["RUN_GLOBAL_SORT"] and cfg["INIT_VCF_LOOKUP"] and cfg["RUN_BINNER"]
    assert cfg["RUN_CNV_CALLER"] and not cfg["RUN_SR_PE_EXTRACTER"]
    assert (cfg["BAM_FILE_REDUCERS"], cfg["TEXT_FILE_REDUCERS"], cfg["REDUCER_TASKS"]) == (800, 100, 24)
    assert cfg["ABBERATION_PENALTY"] == np.float32(0.01) and cfg["TRANSITION_PENALTY"] == np.float32(1.6)


def test_config_properties_syntax(tmp_path, lib):
    p = tmp_path / "c.txt"
    p.write_text("! comment\nBAM_FILE = a b.bam\nVCF_FILE\tx.vcf\nRUN_BINNER=TRUE\nRUN_CNV_CALLER: true \n"
                 "BAM_FILE_REDUCERS=8\nTEXT_FILE_REDUCERS=4\nREDUCER_TASKS= 3 \nABBERATION_PENALTY=-1\nTRANSITION_PENALTY = 0.4 \n")
    cfg = H.config_load(str(p))
    assert cfg["BAM_FILE"] == "a b.bam" and cfg["VCF_FILE"] == "x.vcf"
    assert cfg["RUN_BINNER"] is True
    assert cfg["RUN_CNV_CALLER"] is False          # Boolean.parseBoolean("true ") is false: Properties keeps trailing blanks
    assert cfg["REDUCER_TASKS"] == 3 and cfg["ABBERATION_PENALTY"] == np.float32(-1)
    p.write_text("BAM_FILE_REDUCERS=8 \nTEXT_FILE_REDUCERS=4\nREDUCER_TASKS=3\nABBERATION_PENALTY=1\nTRANSITION_PENALTY=1\n")
    with pytest.raises(H.HcnvError) as e:        # Integer.parseInt("8 ") throws (UserConfig.java:92 has no trim)
        H.config_load(str(p))
    assert e.value.status == L.HCNV_E_PARSE
    with pytest.raises(H.HcnvError) as e:
        H.config_load(str(tmp_path / "missing.txt"))
    assert e.value.status == L.HCNV_E_IO


def test_java_number_formatting_matches_oracle(lib):
    rng = np.random.default_rng(3)
    vals = [0.0, -0.0, 1.0, 100.0, 1e7, 9999999.0, 1e-3, 9.99e-4, 0.1, 1 / 3, 2.0 ** -44, 1e21, 5e-324, float("nan"), float("inf")]
    vals += list(rng.random(500) ** 6) + list(np.ldexp(rng.random(300), rng.integers(-80, 80, 300)))
    for v in vals:
        assert H.format_double(v) == O.double_to_string(v) == Lit.java_double_to_string(v)
        assert H.format_float(v) == O.float_to_string(v)
    assert H.format_float(np.float32(9.808615e-4)) == "9.808615E-4"


def test_text_writers(tmp_path, lib):
    b = H.BinsResult(np.array([0, 100], np.int32), np.array([7, 100], np.int32), np.array([99, 199], np.int32),
                     np.array([17, 0], np.float32), np.array([[0.0625, 0.25, 1 / 3, 0, 1e-5, 2.0 ** -44], [0] * 6]),
                     np.array([1234.0, 12345678.0]), np.array([93, 100], np.int32), np.array([0, 2]))
    p = tmp_path / "part-r-00000"
    H.write_bins_text(str(p), "chr15", b)
    assert p.read_text().splitlines() == [
        "chr15\t0\t7\t99\t17.0\t0.0625\t0.25\t0.3333333333333333\t0.0\t1.0E-5\t5.684341886080802E-14\t1234.0\t93",
        "chr15\t100\t100\t199\t0.0\t0.0\t0.0\t0.0\t0.0\t0.0\t0.0\t1.2345678E7\t100"]
    r = tmp_path / "results.txt"
    H.write_results_text(str(r), "chr15", [20000007], [20000099], [np.float32(-0.6997057)], [2], [2], np.zeros((1, 6), np.float32))
    H.write_results_text(str(r), "chr15", [20000300], [20000399], [np.float32(9.808615e-4)], [2], [2], np.zeros((1, 6), np.float32), append=True)
    assert r.read_text().splitlines() == ["chr15\t20000007\t20000099\t-0.6997057\t2\t2\t0.0\t0.0\t0.0\t0.0\t0.0\t0.0",
                                          "chr15\t20000300\t20000399\t9.808615E-4\t2\t2\t0.0\t0.0\t0.0\t0.0\t0.0\t0.0"]


def test_read_bins_text_parses_like_job_3(tmp_path, lib):
    """bins part file -> arrays: the golden bins lines parse to what the literal restatement's Hmm.init sees, bins come back
    ascending per chromosome even from a shuffled file, and malformed lines are refused."""
    import json
    from oracle import literal as Lit
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "case_2.json")))
    lines = d["bins_lines"]
    other = [ln.replace("chr7\t", "chr8\t", 1) for ln in lines[:50]]
    rng = np.random.default_rng(3)
    mixed = lines + other
    rng.shuffle(mixed)
    p = tmp_path / "part-r-00000"
    p.write_text("\n".join(mixed) + "\n")
    got = H.read_bins_text(str(p))
    assert set(got) == {"chr7", "chr8"}
    c = got["chr7"]
    want = [ln.split("\t") for ln in lines]
    assert c["bin"].tolist() == [int(w[1]) for w in want] and (np.diff(c["bin"]) > 0).all()
    assert c["start"].tolist() == [int(w[2]) for w in want] and c["end"].tolist() == [int(w[3]) for w in want]
    assert c["n"].tolist() == [int(w[12]) for w in want]
    f = lambda s: np.float32(Lit.java_float_value_of(s))
    assert (c["median"].view(np.uint32) == np.array([f(w[4]) for w in want], np.float32).view(np.uint32)).all()
    assert (c["total"].view(np.uint32) == np.array([f(w[11]) for w in want], np.float32).view(np.uint32)).all()
    assert (c["mse"].view(np.uint32) == np.array([[f(x) for x in w[5:11]] for w in want], np.float32).view(np.uint32)).all()
    assert len(got["chr8"]["bin"]) == 50
    bad = tmp_path / "bad"
    bad.write_text("chr1\t0\t1\t99\tnot-a-number\t0\t0\t0\t0\t0\t0\t10.0\t3\n")
    with pytest.raises(H.HcnvError) as e:
        H.read_bins_text(str(bad))
    assert e.value.status == L.HCNV_E_PARSE
