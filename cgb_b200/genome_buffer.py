"""PackedGenome: chromosome / plasmid sequences as 2-bit planes in HBM.

Replaces the reference's sequence store and slicing on the hot path:
``Chromid.sequence`` / ``subsequence`` / ``length`` (cgb/chromid.py:57-60,87-97)
and the per-call ``Seq(seq, alphabet)`` of cgb/pssm_model.py:113.  A genome is a
list of replicons; each is packed on the GPU from its ASCII bytes into

  bases2  2 bits/base (A0 C1 G2 T3), 16 bases per uint32, read as coalesced uint2/uint4
  amb     1 bit/base: byte outside ACGTacgt (what Biopython's _pwm.c turns into NaN)
  lower   1 bit/base: lower-case acgt (only PSSMModel._calculate cares, pssm_model.py:96-100)

at 512-base aligned offsets of one concatenated buffer.  PyTorch only provides
the device memory and the stream; building a PackedGenome performs no device
synchronisation and no cudaMalloc once the caching allocator is warm.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi


def _as_bytes(seq):
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return bytes(seq)
    if isinstance(seq, np.ndarray):
        return seq.astype(np.uint8, copy=False).tobytes()
    return str(seq).encode("ascii")


class HostPackedGenome:
    """One replicon packed ON THE HOST into the kernels' three planes (``cgbk_pack_host``: AVX2,
    host threads; no GPU involved), held in pinned memory when a CUDA runtime is there.

    This is the form a sequence is kept in between runs (and stored in, :meth:`save`): it crosses
    PCIe as 0.25 byte per base (upper-case ACGT only) to 0.5 byte per base instead of 1, and
    ``PSSMModel.regulation_pipeline`` / ``PackedGenome.from_host_packed`` take it without running
    the pack kernel.  Replaces ``Chromid.sequence`` (cgb/chromid.py:57-60) as the sequence store."""

    FILE_MAGIC = "cgbk-packed-genome-1"

    def __init__(self, n_bases, planes, n_amb, n_lower):
        self.n_bases = int(n_bases)
        self.cap_bases = int(_cabi.lib().cgbk_packed_capacity(self.n_bases))
        c16, c32 = self.cap_bases // 16, self.cap_bases // 32
        if planes.dtype != torch.int32 or planes.numel() != c16 + 2 * c32:
            raise ValueError("planes do not match the sequence length")
        self.planes = planes                          # bases2 | amb | lower, one (pinned) int32 tensor
        self.bases2, self.amb, self.lower = planes[:c16], planes[c16:c16 + c32], planes[c16 + c32:]
        self.n_amb, self.n_lower = int(n_amb), int(n_lower)

    @classmethod
    def from_ascii(cls, seq, pin=True, n_threads=0):
        """Pack ``seq`` (str / bytes / uint8 array / uint8 CPU tensor)."""
        if isinstance(seq, torch.Tensor):
            if seq.device.type != "cpu" or seq.dtype != torch.uint8:
                raise ValueError("sequence must be a uint8 CPU tensor")
            src = seq.contiguous().reshape(-1)
        elif isinstance(seq, np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(seq, dtype=np.uint8).reshape(-1))
        else:
            raw = seq if isinstance(seq, (bytes, bytearray)) else str(seq).encode("ascii")
            src = torch.frombuffer(bytearray(raw), dtype=torch.uint8) if len(raw) else torch.empty(0, dtype=torch.uint8)
        n = int(src.numel())
        lib = _cabi.lib()
        cap = int(lib.cgbk_packed_capacity(n))
        planes = torch.empty(cap // 16 + 2 * (cap // 32), dtype=torch.int32)
        if pin and torch.cuda.is_available():
            planes = planes.pin_memory()
        p = planes.data_ptr()
        na, nl = C.c_int64(), C.c_int64()
        _cabi.check(lib.cgbk_pack_host(src.data_ptr() if n else None, n, p, p + cap // 4, p + cap // 4 + cap // 8, cap,
                                       int(n_threads), C.byref(na), C.byref(nl)))
        return cls(n, planes, na.value, nl.value)

    @property
    def upload_bytes(self):
        """Bytes that cross PCIe for this replicon (absent planes are cleared on the device)."""
        return self.cap_bases // 4 + (self.cap_bases // 8 if self.n_amb else 0) + (self.cap_bases // 8 if self.n_lower else 0)

    def save(self, path, name=""):
        """The packed genome file (same container as ``PackedGenome.save``)."""
        with open(path, "wb") as f:
            np.savez(f, magic=np.array(self.FILE_MAGIC), seq_len=np.array([self.n_bases], dtype=np.int64),
                     planes=self.planes.numpy(), names=np.array([name]))

    @classmethod
    def load(cls, path, pin=True):
        """Returns ``(host_packed_genome, name)``; single-replicon files only."""
        with np.load(path, allow_pickle=False) as z:
            if str(z["magic"]) != cls.FILE_MAGIC:
                raise ValueError("%s is not a packed genome file" % path)
            seq_len, planes, names = z["seq_len"], z["planes"], [str(x) for x in z["names"]]
        if len(seq_len) != 1:
            raise ValueError("%s holds %d replicons; load it with PackedGenome.load" % (path, len(seq_len)))
        t = torch.from_numpy(np.ascontiguousarray(planes, dtype=np.int32))
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        cap = int(_cabi.lib().cgbk_packed_capacity(int(seq_len[0])))
        c16, c32 = cap // 16, cap // 32
        n_amb = int(np.unpackbits(planes[c16:c16 + c32].view(np.uint8)).sum())
        n_lower = int(np.unpackbits(planes[c16 + c32:].view(np.uint8)).sum())
        return cls(int(seq_len[0]), t, n_amb, n_lower), names[0]


class PackedGenome:
    def __init__(self, seq_lens, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.CgbkError("PackedGenome needs a CUDA device (there is no CPU path)")
        self.seq_len = np.asarray(seq_lens, dtype=np.int64)
        if self.seq_len.ndim != 1 or len(self.seq_len) == 0:
            raise ValueError("at least one sequence is required")
        A = _cabi.SEQ_ALIGN
        padded = (self.seq_len + A - 1) // A * A
        padded = np.maximum(padded, A)                  # an empty sequence still owns one block
        self.seq_off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
        self.cap_bases = int(self.seq_off[-1] + padded[-1])
        n = len(self.seq_len)
        lib = _cabi.lib()
        # one allocation: bases2 | amb | lower | sequence table.  Every word of the planes is
        # written by the pack kernel (padding included), so nothing is zeroed here.
        c16, c32 = self.cap_bases // 16, self.cap_bases // 32
        n_table = 2 * int(lib.cgbk_seqset_table_ints(n))   # the set's device table (int64) as int32 pairs
        self._storage = torch.empty(c16 + 2 * c32 + n_table, dtype=torch.int32, device=self.device)
        self.bases2 = self._storage[:c16]
        self.amb = self._storage[c16:c16 + c32]
        self.lower = self._storage[c16 + c32:c16 + 2 * c32]
        self._table = self._storage[c16 + 2 * c32:]
        self._struct = _cabi.Genome(self.bases2.data_ptr(), self.amb.data_ptr(), self.lower.data_ptr(),
                                    self.cap_bases)
        self._seqset = C.c_void_p()
        off = (C.c_int64 * n)(*self.seq_off.tolist())
        ln = (C.c_int64 * n)(*self.seq_len.tolist())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(lib.cgbk_seqset_create(off, ln, n, self.device.index or 0, self._table.data_ptr(), stream,
                                           C.byref(self._seqset)))
        self._packed = np.zeros(n, dtype=bool)
        self._work = {}

    # ------------------------------------------------------------------ build
    @classmethod
    def from_sequences(cls, seqs, device=None, pinned=None):
        """Pack ASCII sequences (str / bytes / uint8 arrays).  ``pinned``: optional
        list of pinned uint8 host tensors already holding the bytes (skips the
        staging copy; used by the end-to-end benchmark)."""
        if pinned is None:
            raw = [_as_bytes(s) for s in seqs]
            lens = [len(r) for r in raw]
        else:
            raw = None
            lens = [int(p.numel()) for p in pinned]
        g = cls(lens, device)
        stream = torch.cuda.current_stream(g.device)
        if pinned is None and len(lens) > 1 and g.cap_bases <= (64 << 20):
            # Many host sequences (e.g. the 164 sites of score_self): build the padded image of
            # the whole buffer on the host -- padding as 'A', which packs to the same zero words
            # the per-sequence path writes -- and pack it with ONE copy and ONE launch.
            image = np.full(g.cap_bases, ord("A"), dtype=np.uint8)
            for k, r in enumerate(raw):
                if lens[k]:
                    image[g.seq_off[k]:g.seq_off[k] + lens[k]] = np.frombuffer(r, dtype=np.uint8)
            dev = torch.from_numpy(image).to(g.device, non_blocking=True)
            _cabi.check(_cabi.lib().cgbk_pack_ascii(
                dev.data_ptr(), g.cap_bases, 0, g.bases2.data_ptr(), g.amb.data_ptr(), g.lower.data_ptr(),
                g.cap_bases, stream.cuda_stream))
            dev.record_stream(stream)
            g._packed[:] = True
            return g
        for k in range(len(lens)):
            if pinned is not None:
                host = pinned[k]
            else:
                host = torch.frombuffer(bytearray(raw[k]), dtype=torch.uint8) if lens[k] else \
                    torch.empty(0, dtype=torch.uint8)
            dev = host.to(g.device, non_blocking=True) if lens[k] else torch.empty(0, dtype=torch.uint8,
                                                                                   device=g.device)
            g.pack_into(k, dev, stream)
        return g

    @classmethod
    def from_host_packed(cls, hp, device=None):
        """Device genome of one host-packed replicon: the planes are copied as they are (no pack
        kernel); a plane without a set bit is cleared on the device instead of uploaded."""
        g = cls([hp.n_bases], device)
        c16, c32 = g.cap_bases // 16, g.cap_bases // 32
        g.bases2.copy_(hp.bases2, non_blocking=True)
        if hp.n_amb:
            g.amb.copy_(hp.amb, non_blocking=True)
        else:
            g.amb.zero_()
        if hp.n_lower:
            g.lower.copy_(hp.lower, non_blocking=True)
        else:
            g.lower.zero_()
        g._packed[:] = True
        return g

    def pack_into(self, k, dev_ascii, stream=None):
        """Pack device-resident ASCII bytes as sequence ``k`` (writes the sequence's
        whole aligned block, zero padding included)."""
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        n = int(dev_ascii.numel())
        if n != int(self.seq_len[k]):
            raise ValueError("sequence %d holds %d bytes, expected %d" % (k, n, int(self.seq_len[k])))
        _cabi.check(_cabi.lib().cgbk_pack_ascii(
            dev_ascii.data_ptr() if n else None, n, int(self.seq_off[k]), self.bases2.data_ptr(),
            self.amb.data_ptr(), self.lower.data_ptr(), self.cap_bases, stream.cuda_stream))
        if n:
            dev_ascii.record_stream(stream)       # keep the source alive until the kernel has read it
        self._packed[k] = True

    # ------------------------------------------------------------- accessors
    @property
    def n_seq(self):
        return len(self.seq_len)

    def length(self, k=0):
        """Chromid.length (cgb/chromid.py:87-90)."""
        return int(self.seq_len[k])

    def total_windows(self, m):
        return int(_cabi.lib().cgbk_seqset_total_windows(self._seqset, int(m)))

    def struct_ref(self):
        return C.byref(self._struct)

    @property
    def seqset(self):
        return self._seqset

    def workspace(self, m, cand_cap):
        """Device scratch for the tiled scans of motif length ``m``: one buffer per length, grown to
        the largest candidate capacity asked for so far and handed out whole for smaller requests
        (the library only needs "at least" the size it reports).  A buffer that is replaced may
        still be read by queued kernels: it is kept until the next replacement, by which time the
        stream has moved past them (everything runs on the allocating stream)."""
        key = int(m)
        nbytes = int(_cabi.lib().cgbk_scan_workspace_bytes(self._seqset, int(m), int(cand_cap)))
        w = self._work.get(key)
        if w is None or w.numel() < nbytes:
            self._retired = w
            w = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            if len(self._work) >= 8:                       # many motif lengths over one genome: bound the cache
                self._work.pop(next(iter(self._work)))
            self._work[key] = w
        return w

    # --------------------------------------------------- packed on-disk form
    # a replicon set stored the way the kernels read it (2-bit bases +
    # ambiguity + lower-case planes, 0.5 byte per base) loads with one H2D copy and no pack
    # kernel, instead of parsing and uploading 1 byte per base.
    FILE_MAGIC = "cgbk-packed-genome-1"

    def save(self, path, names=None):
        """Write the packed planes and the sequence table to ``path`` (numpy ``.npz``, not
        compressed: the planes are already 4 bits per base).  Every sequence must have been
        packed.  ``names``: optional accession per sequence."""
        if not bool(self._packed.all()):
            raise ValueError("some sequences have not been packed yet")
        torch.cuda.current_stream(self.device).synchronize()
        planes = self._storage[:self.cap_bases // 16 + 2 * (self.cap_bases // 32)].cpu().numpy()
        with open(path, "wb") as f:
            np.savez(f, magic=np.array(self.FILE_MAGIC), seq_len=self.seq_len, planes=planes,
                     names=np.array(list(names) if names is not None else [""] * self.n_seq))

    @classmethod
    def load(cls, path, device=None):
        """Inverse of :meth:`save`: returns ``(genome, names)``."""
        with np.load(path, allow_pickle=False) as z:
            if str(z["magic"]) != cls.FILE_MAGIC:
                raise ValueError("%s is not a packed genome file" % path)
            seq_len, planes, names = z["seq_len"], z["planes"], [str(x) for x in z["names"]]
        g = cls(seq_len, device)
        want = g.cap_bases // 16 + 2 * (g.cap_bases // 32)
        if planes.dtype != np.int32 or planes.shape != (want,):
            raise ValueError("%s: planes do not match the sequence table" % path)
        g._storage[:want].copy_(torch.from_numpy(planes), non_blocking=False)
        g._packed[:] = True
        return g, names

    def __del__(self):
        try:
            if getattr(self, "_seqset", None) is not None and self._seqset.value:
                _cabi.lib().cgbk_seqset_destroy(self._seqset)
                self._seqset = C.c_void_p()
        except Exception:
            pass
