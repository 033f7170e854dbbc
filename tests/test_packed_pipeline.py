"""The host packer (cgbk_pack_host) and the packed-input pipeline (cgbk_regulation_pipeline_packed).

CPU: the planes against a numpy restatement of their definition (include/cgbk.h), ragged lengths,
every byte class, the packed genome file.  GPU: the host planes bit for bit against the device pack
kernel, and the packed pipeline against the ASCII pipeline and the C oracle (chromid.py:57-60 is
what the packed form replaces; the numbers that come out must not depend on the form)."""
import os

import numpy as np
import pytest
import torch

from cgb_b200 import GeneTable, HostPackedGenome, PackedGenome, PSSMModel, SiteCollection, _cabi
from oracle import c_oracle, np_oracle as O

ALPHABET = np.frombuffer(b"ACGTacgtNnRYKM-*xZ\x00\xff", dtype=np.uint8)


def _mixed(rng, n, p_other=0.12):
    p = np.r_[np.full(8, (1 - p_other) / 8), np.full(len(ALPHABET) - 8, p_other / (len(ALPHABET) - 8))]
    return ALPHABET[rng.choice(len(ALPHABET), size=n, p=p)]


def _planes_by_definition(seq, cap):
    pad = np.full(cap, ord("A"), dtype=np.uint8)
    pad[:len(seq)] = seq
    up = pad & 0xDF
    valid = (up == 65) | (up == 67) | (up == 71) | (up == 84)
    code = np.zeros(cap, dtype=np.uint32)
    code[up == 67], code[up == 71], code[up == 84] = 1, 2, 3
    code[~valid] = 0
    b2 = np.zeros(cap // 16, dtype=np.uint32)
    for k in range(16):
        b2 |= code[k::16] << np.uint32(2 * k)
    amb = np.packbits((~valid).astype(np.uint8), bitorder="little").view(np.uint32)
    low = np.packbits((valid & ((pad & 0x20) != 0)).astype(np.uint8), bitorder="little").view(np.uint32)
    return b2, amb, low


@pytest.mark.parametrize("n", [0, 1, 15, 31, 32, 33, 511, 512, 513, 4097, 300_001])
def test_host_packer_matches_the_plane_definition(n):
    rng = np.random.default_rng(n + 1)
    seq = _mixed(rng, n)
    hp = HostPackedGenome.from_ascii(seq, pin=False)
    assert hp.n_bases == n and hp.cap_bases % 512 == 0 and hp.cap_bases >= max(n, 512)
    b2, amb, low = _planes_by_definition(seq, hp.cap_bases)
    assert np.array_equal(hp.bases2.numpy().view(np.uint32), b2)
    assert np.array_equal(hp.amb.numpy().view(np.uint32), amb)
    assert np.array_equal(hp.lower.numpy().view(np.uint32), low)
    assert hp.n_amb == int(np.unpackbits(amb.view(np.uint8)).sum())
    assert hp.n_lower == int(np.unpackbits(low.view(np.uint8)).sum())


def test_host_packer_threads_and_inputs_agree():
    rng = np.random.default_rng(5)
    seq = _mixed(rng, 4_500_123, p_other=0.01)             # large enough for the threaded split
    one = HostPackedGenome.from_ascii(seq, pin=False, n_threads=1)
    many = HostPackedGenome.from_ascii(seq.tobytes(), pin=False, n_threads=4)
    assert torch.equal(one.planes, many.planes) and (one.n_amb, one.n_lower) == (many.n_amb, many.n_lower)
    t = HostPackedGenome.from_ascii(torch.from_numpy(seq), pin=False)
    assert torch.equal(one.planes, t.planes)
    clean = HostPackedGenome.from_ascii("ACGT" * 1000, pin=False)
    assert clean.n_amb == 0 and clean.n_lower == 0 and clean.upload_bytes == clean.cap_bases // 4


def test_packed_genome_file_round_trip(tmp_path):
    rng = np.random.default_rng(6)
    hp = HostPackedGenome.from_ascii(_mixed(rng, 70_001), pin=False)
    path = os.path.join(tmp_path, "g.npz")
    hp.save(path, "NC_TEST.1")
    back, name = HostPackedGenome.load(path, pin=False)
    assert name == "NC_TEST.1" and back.n_bases == hp.n_bases and torch.equal(back.planes, hp.planes)
    assert (back.n_amb, back.n_lower) == (hp.n_amb, hp.n_lower)


def test_pack_host_rejects_bad_capacity():
    import ctypes as C
    lib = _cabi.lib()
    buf = np.zeros(64, dtype=np.int32)
    seq = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
    p = buf.ctypes.data
    assert lib.cgbk_pack_host(seq.ctypes.data, 40, p, p + 32, p + 64, 100, 1, None, None) != 0
    assert b"cap_bases" in lib.cgbk_last_error()
    assert lib.cgbk_pack_host(seq.ctypes.data, 40, None, p + 32, p + 64, 512, 1, None, None) != 0


# ----------------------------------------------------------------------------- GPU
def _collections(lexa):
    return [SiteCollection(mo["sites"], None, mo["name"]) for mo in lexa["motifs"]]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 700, 100_003])
def test_host_planes_equal_device_pack_kernel(n):
    rng = np.random.default_rng(n)
    seq = _mixed(rng, n)
    hp = HostPackedGenome.from_ascii(seq)
    g = PackedGenome.from_sequences([seq])
    torch.cuda.synchronize()
    assert torch.equal(g.bases2.cpu(), hp.bases2) and torch.equal(g.amb.cpu(), hp.amb)
    assert torch.equal(g.lower.cpu(), hp.lower)
    g2 = PackedGenome.from_host_packed(hp)
    torch.cuda.synchronize()
    assert torch.equal(g2._storage[:hp.planes.numel()].cpu(), hp.planes)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,chunks", [("clean", 0), ("amb", 0), ("mixed", 0), ("mixed", 3), ("mixed", 4)])
def test_packed_pipeline_equals_ascii_pipeline_and_oracle(lexa, lexa_weights, kind, chunks, monkeypatch):
    if chunks:
        monkeypatch.setenv("CGBK_PIPE_CHUNKS", str(chunks))
    n = 600_037
    seq = np.frombuffer(O.synthetic_genome(n, 31), dtype=np.uint8).copy()
    if kind != "clean":
        for cut in (n // 2 // 512 * 512, int(0.8 * n) // 512 * 512, int(0.35 * n) // 512 * 512):
            seq[cut - 5:cut + 11] = ord("N")
        seq[7000:7003] = ord("R")
    if kind == "mixed":
        seq[100:140] = np.frombuffer(b"acgtacgtacgtacgtacgtacgtacgtacgtacgtacgt", dtype=np.uint8)
        seq[300_000:300_400] |= 0x20
    hp = HostPackedGenome.from_ascii(seq)
    assert (hp.n_amb > 0) == (kind != "clean") and (hp.n_lower > 0) == (kind == "mixed")
    n_genes = 300
    step = n // n_genes
    start = np.arange(n_genes, dtype=np.int64) * step + step // 3
    gt = GeneTable(start, start + 900, np.where(np.arange(n_genes) % 2 == 0, 1, -1), n)
    ps, pe = gt.promoter_region_location(300, 50)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    post_p, stats_p, hist_p = M.regulation_pipeline(hp, ps, pe, prior, alpha, n_bins=256)
    M2 = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    post_a, stats_a, hist_a = M2.regulation_pipeline(seq, ps, pe, prior, alpha, n_bins=256)
    # the two forms run the same kernels on the same planes: identical, not merely close
    assert np.array_equal(hist_p, hist_a) and np.array_equal(stats_p, stats_a) and np.array_equal(post_p, post_a)
    # and the oracle (binding_model.py:94-113 on the reference's own score_seq)
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], lexa_weights["site_count"])
    raw = seq.tobytes()
    _, _, b = c_oracle.score_seq(om, raw)
    O.build_bayesian_estimator(om, b)
    assert stats_p[_cabi.BG_N] == len(b) == hist_p.sum()
    assert abs(stats_p[_cabi.BG_MEAN] - om._mu_bg) < 1e-6 and abs(stats_p[_cabi.BG_STD] - om._std_bg) < 1e-6
    s2, e2 = np.clip(ps, 0, n), np.clip(pe, 0, n)
    exp = np.array([c_oracle.binding_probability(om, raw[int(a):int(b_)], om._mu_m, om._std_m, om._mu_bg,
                                                 om._std_bg, prior, alpha) for a, b_ in zip(s2, e2)])
    assert np.max(np.abs(post_p - exp)) < 1e-6


@pytest.mark.gpu
def test_packed_pipeline_c_call_with_pageable_buffers(lexa, lexa_weights):
    """The plain (non-struct) C entry point, NULL flag planes, pageable result buffers."""
    import ctypes as C
    n = 200_000
    seq = np.frombuffer(O.synthetic_genome(n, 32), dtype=np.uint8).copy()
    hp = HostPackedGenome.from_ascii(seq, pin=False)
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    ps = np.arange(10, dtype=np.int64) * 15000 + 100
    pe = ps + 350
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    ref_post, ref_stats, ref_hist = M.regulation_pipeline(seq, ps, pe, prior, alpha, n_bins=64)
    lib = _cabi.lib()
    dev = torch.device("cuda", 0)
    scratch = torch.empty(int(lib.cgbk_pipeline_scratch_bytes(n, 10, M.length, 64)), dtype=torch.uint8, device=dev)
    staging = np.zeros(int(lib.cgbk_pipeline_staging_bytes(10)), dtype=np.uint8)
    post, stats, hist = np.zeros(10), np.zeros(_cabi.BG_COUNT), np.zeros(64, dtype=np.uint64)
    lo, hi = M._score_bounds
    _cabi.check(lib.cgbk_regulation_pipeline_packed(
        M._model(), hp.bases2.data_ptr(), None, None, n, ps.ctypes.data, pe.ctypes.data, 10,
        float(M._mu_m), float(M._std_m), prior, alpha, float(lo), float(hi), 64,
        scratch.data_ptr(), scratch.numel(), staging.ctypes.data, staging.size,
        post.ctypes.data, stats.ctypes.data, hist.ctypes.data, torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(post, ref_post) and np.array_equal(hist, ref_hist) and np.array_equal(stats, ref_stats)
