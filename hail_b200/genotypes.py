"""The device genotype store: 2-bit packed calls resident in HBM (ingest row of SURVEY.md 8).

Replaces, for the GWAS case x = GT.n_alt_alleles(), the reference's per-entry region values
(`array<struct{x: float64}>`, 16 B per entry; hail/hail/src/is/hail/types/physical/PCanonicalArray.scala:48-117)
with 0.25 B per call.  Format authority for the PLINK side: hail/hail/src/is/hail/io/plink/LoadPlink.scala:37-38,
240-251 (magic + size check), :475-481 and :525 (codes and bit order).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])  # LoadPlink.scala:37-38 ; ExportPlink.scala:10


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class PackedGenotypes:
    """[n_variants, stride] uint8 on one CUDA device, rows in the lrr_b200 store layout (include/lrr_b200.h)."""

    def __init__(self, data: torch.Tensor, n_variants: int, n_samples: int, row_flags: torch.Tensor = None):
        assert data.is_cuda and data.dtype == torch.uint8 and data.dim() == 2
        self.data = data
        # the "missing mask" side array: uint8 [M], 1 iff the row has a missing call (None = unknown)
        self.row_flags = row_flags
        self.n_variants = int(n_variants)
        self.n_samples = int(n_samples)
        self.stride = int(data.shape[1])
        assert data.shape[0] == self.n_variants
        assert self.stride == packed_stride(self.n_samples)

    @property
    def device(self):
        return self.data.device

    @property
    def nbytes(self):
        return self.n_variants * self.stride

    # ---- constructors -----------------------------------------------------------------------
    @classmethod
    def empty(cls, n_variants, n_samples, device=0):
        dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        stride = packed_stride(n_samples)
        return cls(torch.empty((n_variants, stride), dtype=torch.uint8, device=dev), n_variants, n_samples,
                   torch.zeros(n_variants, dtype=torch.uint8, device=dev))

    @classmethod
    def from_bed_rows(cls, rows, n_samples, device=0, chunk_variants=1 << 16):
        """rows: uint8 [M, >= ceil(N/4)] PLINK SNP-major body (host numpy or torch)."""
        rows_t = torch.as_tensor(np.ascontiguousarray(rows) if isinstance(rows, np.ndarray) else rows)
        assert rows_t.dtype == torch.uint8 and rows_t.dim() == 2
        M, bed_stride = rows_t.shape
        if bed_stride < (n_samples + 3) // 4:
            raise ValueError("bed row shorter than ceil(n_samples/4) bytes (LoadPlink.scala:240-251)")
        out = cls.empty(M, n_samples, device)
        ctx = _lib.context(out.device.index)
        with torch.cuda.device(out.device):
            for lo in range(0, M, chunk_variants):
                hi = min(M, lo + chunk_variants)
                d_in = rows_t[lo:hi].to(out.device, non_blocking=True).contiguous()
                ctx.check(ctx.lib.lrr_pack_bed(ctx.handle, d_in.data_ptr(), hi - lo, bed_stride, n_samples,
                                               out.data[lo:hi].data_ptr(), out.stride,
                                               out.row_flags[lo:hi].data_ptr(), _stream_ptr(out.device)))
        return out

    @classmethod
    def from_bed_file(cls, path, n_samples, n_variants, device=0):
        raw = np.fromfile(path, dtype=np.uint8)
        stride = (n_samples + 3) // 4
        if raw.size < 3 or bytes(raw[:3]) != BED_MAGIC:
            raise ValueError(f"{path}: not a SNP-major PLINK .bed (bad magic)")
        if raw.size != 3 + n_variants * stride:
            raise ValueError(f"{path}: size {raw.size} != 3 + n_variants*ceil(n_samples/4) = {3 + n_variants * stride}")
        return cls.from_bed_rows(raw[3:].reshape(n_variants, stride), n_samples, device)

    @classmethod
    def from_dosage(cls, dosage, device=0, chunk_variants=1 << 14):
        """dosage: int8 [M, N] with 0/1/2 and anything else (e.g. -1) = missing; float arrays with NaN accepted."""
        d = np.asarray(dosage)
        if d.dtype.kind == "f":
            bad = ~np.isin(d, (0.0, 1.0, 2.0))
            d = np.where(bad, -1, d).astype(np.int8)
        d = np.ascontiguousarray(d, dtype=np.int8)
        M, N = d.shape
        out = cls.empty(M, N, device)
        ctx = _lib.context(out.device.index)
        with torch.cuda.device(out.device):
            for lo in range(0, M, chunk_variants):
                hi = min(M, lo + chunk_variants)
                d_in = torch.from_numpy(d[lo:hi]).to(out.device)
                ctx.check(ctx.lib.lrr_pack_dosage_i8(ctx.handle, d_in.data_ptr(), hi - lo, N,
                                                     out.data[lo:hi].data_ptr(), out.stride,
                                                     out.row_flags[lo:hi].data_ptr(), _stream_ptr(out.device)))
        return out

    # ---- views / export ---------------------------------------------------------------------
    def rows(self, lo, hi):
        lo, hi = int(lo), int(hi)
        return PackedGenotypes(self.data[lo:hi], hi - lo, self.n_samples,
                               None if self.row_flags is None else self.row_flags[lo:hi])

    def flags_ptr(self, lo=0):
        return None if self.row_flags is None else self.row_flags[lo:].data_ptr()

    def to_dosage(self) -> np.ndarray:
        """int8 [M, N], missing = -1."""
        ctx = _lib.context(self.device.index)
        out = torch.empty((self.n_variants, self.n_samples), dtype=torch.int8, device=self.device)
        with torch.cuda.device(self.device):
            ctx.check(ctx.lib.lrr_unpack_dosage_i8(ctx.handle, self.data.data_ptr(), self.stride, self.n_variants,
                                                   self.n_samples, out.data_ptr(), _stream_ptr(self.device)))
        return out.cpu().numpy()


class HostBedGenotypes:
    """A SNP-major PLINK .bed body kept in page-locked HOST memory: uint8 [n_variants, ceil(n_samples/4)].

    The out-of-core form of the entry matrix: `linear_regression_rows` streams it through the device block by
    block (lrr_stream_*), the way the reference's per-partition loop consumes rows as LoadPlink decodes them
    (LinearRegression.scala:95; io/plink/LoadPlink.scala:470-530).  Nothing is resident on the device between calls.
    """

    def __init__(self, rows, n_samples: int, device=0):
        t = torch.as_tensor(np.ascontiguousarray(rows) if isinstance(rows, np.ndarray) else rows)
        assert t.dtype == torch.uint8 and t.dim() == 2
        if t.shape[1] < (n_samples + 3) // 4:
            raise ValueError("bed row shorter than ceil(n_samples/4) bytes (LoadPlink.scala:240-251)")
        if t.is_cuda:
            raise ValueError("HostBedGenotypes holds host memory; use PackedGenotypes.from_bed_rows for device data")
        if not t.is_pinned():  # one copy into page-locked memory so that the block copies run at PCIe rate
            p = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
            p.copy_(t)
            t = p
        self.rows = t.contiguous()
        self.n_variants, self.bed_stride = int(t.shape[0]), int(t.shape[1])
        self.n_samples = int(n_samples)
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device

    @property
    def nbytes(self):
        return self.n_variants * self.bed_stride

    @classmethod
    def from_bed_file(cls, path, n_samples, n_variants, device=0):
        """Read `path` (magic 6c 1b 01, SNP-major) straight into page-locked memory."""
        stride = (n_samples + 3) // 4
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            if size < 3 or f.read(3) != BED_MAGIC:
                raise ValueError(f"{path}: not a SNP-major PLINK .bed (bad magic)")
            if size != 3 + n_variants * stride:
                raise ValueError(f"{path}: size {size} != 3 + n_variants*ceil(n_samples/4) = {3 + n_variants * stride}")
            t = torch.empty((n_variants, stride), dtype=torch.uint8, pin_memory=True)
            if n_variants:
                got = f.readinto(memoryview(t.numpy()).cast("B"))
                assert got == n_variants * stride
        return cls(t, n_samples, device)

    def to_device(self, chunk_variants=1 << 14) -> PackedGenotypes:
        """The resident 2-bit store of the same calls."""
        return PackedGenotypes.from_bed_rows(self.rows, self.n_samples, self.device, chunk_variants=chunk_variants)

    def to_dosage(self) -> np.ndarray:
        return self.to_device().to_dosage()


class DenseDosage:
    """A dense float64 entry field on the device: [n_variants, n_samples], NaN = missing.

    The general form of `x` in `linear_regression_rows` (any entry-indexed float64 expression, statgen.py:229, 391):
    PL / GP dosages, imputed dosages, ...  8 bytes per entry instead of 0.25: use PackedGenotypes for hard calls.
    """

    def __init__(self, values, device=0):
        dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        t = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64) if isinstance(values, np.ndarray) else values)
        assert t.dim() == 2
        self.data = t.to(device=dev, dtype=torch.float64).contiguous()
        self.n_variants, self.n_samples = int(t.shape[0]), int(t.shape[1])

    @property
    def device(self):
        return self.data.device

    @property
    def nbytes(self):
        return self.data.numel() * 8

    def to_dosage(self) -> np.ndarray:
        return self.data.cpu().numpy()


class CompactDosage:
    """A dosage entry field stored as one uint16 per entry on the device: value = q * scale, q = 0xFFFF = missing
    (include/lrr_b200.h lrr_run_dense_u16).  The compact form of `DenseDosage` for imputed data: BGEN's 8-bit genotype
    probabilities give dosages that are exact multiples of 1/255 (`scale=1/255`, the default when every defined value
    is one); otherwise `scale = 2/65534` stores any dosage in [0, 2] to 1.5e-5.  2 bytes of HBM per entry instead of 8.
    The regression sees the DEQUANTISED values: `to_dosage()` returns exactly them."""

    MISSING = 0xFFFF

    def __init__(self, values, scale=None, device=0):
        dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        v = np.asarray(values, dtype=np.float64)
        assert v.ndim == 2
        miss = np.isnan(v)
        if scale is None:
            q255 = np.where(miss, 0.0, v) * 255.0
            scale = 1.0 / 255.0 if np.all(np.abs(q255 - np.rint(q255)) < 1e-9) else 2.0 / 65534.0
        q = np.rint(np.where(miss, 0.0, v) / scale)
        if q.min(initial=0.0) < 0 or q.max(initial=0.0) > 65534:
            raise ValueError("CompactDosage: values must lie in [0, 65534 * scale]")
        q = np.where(miss, self.MISSING, q).astype(np.uint16)
        M, N = q.shape
        ld = (N + 7) // 8 * 8
        buf = np.zeros((M, ld), dtype=np.uint16)
        buf[:, :N] = q
        # (torch has no uint16 arithmetic, but it can hold the bytes: int16 view of the same bits)
        self.data = torch.from_numpy(buf.view(np.int16)).to(dev).contiguous()
        self.scale = float(scale)
        self.ld = ld
        self.n_variants, self.n_samples = int(M), int(N)

    @classmethod
    def from_q(cls, q, scale, device=0):
        """From the stored integers themselves: q uint16 [M, N] (0xFFFF = missing), value = q * scale -- what a BGEN file's
        8-bit probabilities give directly (q = P(het) + 2 P(hom alt) in units of 1/255, `impex.import_bgen`)."""
        dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        q = np.asarray(q)
        assert q.ndim == 2 and q.dtype == np.uint16
        self = cls.__new__(cls)
        M, N = q.shape
        ld = (N + 7) // 8 * 8
        buf = np.zeros((M, ld), dtype=np.uint16)
        buf[:, :N] = q
        self.data = torch.from_numpy(buf.view(np.int16)).to(dev).contiguous()
        self.scale = float(scale)
        self.ld = ld
        self.n_variants, self.n_samples = int(M), int(N)
        return self

    @property
    def device(self):
        return self.data.device

    @property
    def nbytes(self):
        return self.data.numel() * 2

    def to_dosage(self) -> np.ndarray:
        q = self.data.cpu().numpy().view(np.uint16)[:, : self.n_samples].astype(np.float64)
        return np.where(q == self.MISSING, np.nan, q * self.scale)


def packed_stride(n_samples: int) -> int:
    return int(_lib.load().lrr_packed_stride(int(n_samples)))


def read_plink_shape(bim_path, fam_path):
    with open(bim_path) as f:
        n_variants = sum(1 for line in f if line.strip())
    with open(fam_path) as f:
        n_samples = sum(1 for line in f if line.strip())
    return n_variants, n_samples
