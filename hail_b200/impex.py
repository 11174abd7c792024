"""PLINK ingest / export for the B200 path: .bed/.bim/.fam <-> MatrixTable with packed GT;
BGEN ingest: 8-bit genotype probabilities -> the compact uint16 dosage store.

Reference behaviour followed:
  * `hl.import_plink` / `hl.import_fam`    hail/python/hail/methods/impex.py:2505 (defaults: a2_reference=True, missing='NA',
                                           delimiter='\\\\s+', quant_pheno=False); driver checks and messages
                                           hail/hail/src/is/hail/io/plink/LoadPlink.scala:40-82 (bim), :102-186 (fam),
                                           :225-251 (magic, SNP-major, file size), :475-481 + :525 (entry decode);
                                           rows are SORTED by (locus, alleles) (:79-81), each keeping its .bed row index
  * `hl.export_plink`                      impex.py:324-470 (defaults, white-space check),
                                           hail/hail/src/is/hail/expr/ir/MatrixWriter.scala:2110-2285 (bytes written)
  * `hl.import_bgen`                       impex.py:1100-1290 (entry_fields, sample_file); the decode and its fatal conditions
                                           hail/hail/src/is/hail/io/bgen/StagedBGENReader.scala:120-520 (layout 2, biallelic,
                                           diploid, unphased, 8 bits per probability; dosage = (d1 + 2 d2) / 255 with
                                           d2 = 255 - d0 - d1, :350-353; an entry is missing when bit 7 of its ploidy byte is
                                           set, :498), header / sample block hail/hail/src/is/hail/io/bgen/LoadBgen.scala
The genotype bytes never pass through Python objects: they are read into page-locked memory (`resident=False`: the
out-of-core form streamed by lrr_stream_*) or packed into the device store (`resident=True`).
"""
from __future__ import annotations

import os
import re

import numpy as np

from .genotypes import BED_MAGIC, CompactDosage, HostBedGenotypes, PackedGenotypes
from .matrixtable import MatrixTable
from .statgen import FatalError

_NUMERIC = re.compile(r"^-?(?:\d+|\d*\.\d+)(?:[eE]-?\d+)?$")   # LoadPlink.scala:84-85
_GRCH37_CONTIGS = [str(i) for i in range(1, 23)] + ["X", "Y", "MT"]

# a2_reference=False swaps the homozygous codes of a .bed byte (00 <-> 11 per 2-bit field; 01 missing, 10 het stay)
_SWAP_HOM = np.array([b ^ (((~(b ^ (b >> 1))) & 0x55) * 3) for b in range(256)], dtype=np.uint8)


def _plural(n, w):
    return w if n == 1 else w + "s"


def import_fam(path, quant_pheno=False, delimiter=r"\\s+", missing="NA"):
    """Columns of a .fam file as a dict of arrays (LoadPlink.parseFam, LoadPlink.scala:102-186).

    `id` str, `fam_id` / `pat_id` / `mat_id` str or None ('0' = missing), `is_female` float (1.0 / 0.0 / NaN),
    `is_case` float (1.0 / 0.0 / NaN) or `quant_pheno` float64 (NaN = missing).
    """
    delim = delimiter.replace("\\\\", "\\")
    ids, fam_id, pat, mat, sex, pheno = [], [], [], [], [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            split = re.split(delim, line)
            if len(split) != 6:
                raise FatalError(f"expected 6 fields, but found {len(split)}")
            fam, kid, dad, mom, is_female, ph = split
            fam_id.append(fam if fam != "0" else None)
            pat.append(dad if dad != "0" else None)
            mat.append(mom if mom != "0" else None)
            if is_female in (missing, "-9", "0"):
                sex.append(np.nan)
            elif is_female == "1":
                sex.append(0.0)
            elif is_female == "2":
                sex.append(1.0)
            else:
                raise FatalError(f"Invalid sex: '{is_female}'. Male is '1', female is '2', unknown is '0'")
            if quant_pheno:
                if ph == missing:
                    pheno.append(np.nan)
                elif ph == "-9":   # a valid quantitative phenotype in Hail (unlike PLINK), LoadPlink.scala:147-155
                    pheno.append(-9.0)
                elif _NUMERIC.match(ph):
                    pheno.append(float(ph))
                else:
                    raise FatalError(f"Invalid quantitative phenotype: '{ph}'. Value must be numeric or '{missing}'")
            else:
                if ph == "1":
                    pheno.append(0.0)
                elif ph == "2":
                    pheno.append(1.0)
                elif ph in (missing, "0", "-9", "N/A"):
                    pheno.append(np.nan)
                elif _NUMERIC.match(ph):
                    raise FatalError(f"Invalid case-control phenotype: '{ph}'. Control is '1', case is '2', missing is "
                                     f"'0', '-9', '{missing}', or non-numeric.")
                else:
                    pheno.append(np.nan)
            ids.append(kid)
    if not ids:
        raise FatalError("Empty FAM file")
    out = {"id": np.array(ids, dtype=object), "fam_id": np.array(fam_id, dtype=object),
           "pat_id": np.array(pat, dtype=object), "mat_id": np.array(mat, dtype=object),
           "is_female": np.array(sex, dtype=np.float64)}
    out["quant_pheno" if quant_pheno else "is_case"] = np.array(pheno, dtype=np.float64)
    return out


def _parse_bim(path, a2_reference, contig_recoding, reference_genome, skip_invalid_loci):
    """-> (n_total_lines, kept file indices, contig, position, alleles, rsid, cm) in FILE order (LoadPlink.scala:40-82)."""
    rank = {c: i for i, c in enumerate(_GRCH37_CONTIGS)} if reference_genome is not None else None
    idx, contig, pos, alleles, rsid, cm = [], [], [], [], [], []
    n = 0
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            r = re.split(r"\s+", line)
            if len(r) != 6:
                raise FatalError(f"Invalid .bim line.  Expected 6 fields, found {len(r)} {_plural(len(r), 'field')}")
            c = contig_recoding.get(r[0], r[0])
            p = int(r[3])
            valid = rank is None or (c in rank and p >= 1)
            if not valid and not skip_invalid_loci:
                raise FatalError(f"Invalid locus '{c}:{p}' found. Contig '{c}' is not in the reference genome "
                                 f"'{reference_genome}'." if c not in rank else
                                 f"Invalid locus '{c}:{p}' found. Position '{p}' is not within the range [1-...] "
                                 f"for reference genome '{reference_genome}'.")
            if valid:
                idx.append(n)
                contig.append(c)
                pos.append(p)
                alleles.append((r[5], r[4]) if a2_reference else (r[4], r[5]))
                rsid.append(r[1])
                cm.append(float(r[2]))
            n += 1
    return n, np.array(idx, dtype=np.int64), contig, np.array(pos, dtype=np.int64), alleles, rsid, np.array(cm), rank


def import_plink(bed, bim, fam, n_partitions=None, block_size=None, min_partitions=None, missing="NA",
                 delimiter=r"\\s+", quant_pheno=False, a2_reference=True, reference_genome="default",
                 contig_recoding=None, skip_invalid_loci=False, *, device=0, resident=True) -> MatrixTable:
    """`hl.import_plink` (impex.py:2505).  Row fields `locus`, `alleles`, `rsid`, `cm_position` (key locus, alleles,
    rows sorted by key as LoadPlink.scala:79-81 does); column fields `s` (key), `fam_id`, `pat_id`, `mat_id`,
    `is_female`, and `is_case` or `quant_pheno`; entry field `GT`.

    `resident=True` packs the calls into HBM (2 bit each); `resident=False` keeps the .bed body in page-locked host
    memory and `linear_regression_rows` streams it through the device.  `n_partitions` / `block_size` /
    `min_partitions` are accepted and inert (there are no partitions here).
    """
    if reference_genome == "default":
        reference_genome = "GRCh37"
    if reference_genome not in (None, "GRCh37"):
        raise NotImplementedError("import_plink: only reference_genome='GRCh37' (contig order 1-22, X, Y, MT) or None")
    cols_fam = import_fam(fam, quant_pheno=quant_pheno, delimiter=delimiter, missing=missing)
    n_samples = len(cols_fam["id"])
    if n_samples <= 0:
        raise FatalError("FAM file does not contain any samples")
    n_total, idx, contig, pos, alleles, rsid, cm, rank = _parse_bim(bim, a2_reference, contig_recoding or {},
                                                                  reference_genome, skip_invalid_loci)
    if n_total <= 0:
        raise FatalError("BIM file does not contain any variants")
    with open(bed, "rb") as f:
        head = f.read(3)
    if len(head) < 2 or head[0] != 108 or head[1] != 27:
        raise FatalError("First two bytes of BED file do not match PLINK magic numbers 108 & 27")
    if len(head) < 3 or head[2] == 0:
        raise FatalError("BED file is in individual major mode. First use plink with --make-bed to convert file to snp "
                         "major mode before using Hail")
    stride = (n_samples + 3) // 4
    if os.path.getsize(bed) != 3 + n_total * stride:
        raise FatalError("BED file size does not match expected number of bytes based on BIM and FAM files")

    # key order: (locus, alleles); locus order = contig rank in the reference genome (or the contig string), position
    keys = [((rank[c] if rank is not None else c), int(p), a) for c, p, a in zip(contig, pos, alleles)]
    order = np.array(sorted(range(len(keys)), key=keys.__getitem__), dtype=np.int64)
    file_rows = idx[order]                                     # .bed row of every MatrixTable row
    host = HostBedGenotypes.from_bed_file(bed, n_samples, n_total, device)
    identity = len(file_rows) == n_total and np.array_equal(file_rows, np.arange(n_total))
    if not a2_reference:
        import torch
        host.rows.copy_(torch.from_numpy(_SWAP_HOM[host.rows.numpy()]))
        pad = (-n_samples) % 4     # the pad calls of the last byte of a row are 00 in the file and must stay 00
        if pad:
            host.rows[:, -1] &= 0xFF >> (2 * pad)
    if resident:
        gt = host.to_device()
        if not identity:
            import torch
            sel = torch.from_numpy(file_rows).to(gt.device)
            gt = PackedGenotypes(gt.data.index_select(0, sel).contiguous(), len(file_rows), n_samples,
                                 gt.row_flags.index_select(0, sel).contiguous())
    else:
        gt = host if identity else HostBedGenotypes(host.rows[file_rows], n_samples, device)
    take = lambda seq: [seq[i] for i in order]
    rows = {
        "locus": np.array([(c, int(p)) for c, p in zip(take(contig), pos[order])] + [None], dtype=object)[:-1],
        "alleles": np.array(take(alleles) + [None], dtype=object)[:-1],
        "rsid": np.array(take(rsid), dtype=object),
        "cm_position": cm[order],
    }
    cols = {"s": cols_fam["id"], "fam_id": cols_fam["fam_id"], "pat_id": cols_fam["pat_id"], "mat_id": cols_fam["mat_id"],
            "is_female": cols_fam["is_female"]}
    key = "quant_pheno" if quant_pheno else "is_case"
    cols[key] = cols_fam[key]
    return MatrixTable(gt, rows=rows, cols=cols, row_key=("locus", "alleles"), col_key=("s",))


def _strings(v, n, default, what):
    if v is None:
        return [default] * n
    vals = v.values if hasattr(v, "values") and not isinstance(v, np.ndarray) else v
    out = []
    for x in np.asarray(vals, dtype=object):
        out.append(default if x is None or (isinstance(x, float) and np.isnan(x)) else str(x))
    if len(out) != n:
        raise ValueError(f"export_plink/{what}: expected {n} values, found {len(out)}")
    return out


def export_plink(dataset: MatrixTable, output, call=None, fam_id=None, ind_id=None, pat_id=None, mat_id=None,
                 is_female=None, pheno=None, varid=None, cm_position=None):
    """`hl.export_plink` (impex.py:324-470): writes `output`.bed / .bim / .fam.

    Defaults as the reference's: fam_id / pat_id / mat_id '0', is_female '0' ('2' female, '1' male), pheno 'NA'
    ('2' / '1' for booleans), varid 'contig:position:ref:alt', cm_position 0.0; A1 = alleles[1], A2 = alleles[0]
    (MatrixWriter.scala:2236-2262).  IDs containing white space raise TypeError (impex.py:452-463).
    """
    n, m = dataset.count_cols(), dataset.count_rows()
    if ind_id is None:
        if len(dataset.col_key) != 1:
            raise ValueError("export_plink: 'ind_id' is required unless the column key is one string field")
        ind_id = dataset._cols[dataset.col_key[0]]
    fam_cols = {"fam_id": _strings(fam_id, n, "0", "fam_id"), "ind_id": _strings(ind_id, n, "0", "ind_id"),
                "pat_id": _strings(pat_id, n, "0", "pat_id"), "mat_id": _strings(mat_id, n, "0", "mat_id")}
    errors = []
    for name in ("ind_id", "fam_id", "pat_id", "mat_id"):
        bad = [v for v in fam_cols[name] if re.search(r"\s+", v)]
        if bad:
            errors.append(f"expr '{name}' has spaces in the following values:\n")
            errors.extend(f"  {v}\n" for v in bad)
    if errors:
        raise TypeError("\n".join(errors))

    def values_of(v):
        return np.asarray(v.values if hasattr(v, "values") and not isinstance(v, np.ndarray) else v)

    def is_missing(x):
        return x is None or (isinstance(x, (float, np.floating)) and np.isnan(x))

    sex = ["0"] * n if is_female is None else ["0" if is_missing(x) else ("2" if bool(x) else "1")
                                               for x in values_of(is_female).astype(object)]
    if pheno is None:
        ph = ["NA"] * n
    else:
        pv = values_of(pheno)
        if pv.dtype == bool:
            ph = ["2" if x else "1" for x in pv]
        else:
            ph = ["NA" if is_missing(x) else repr(float(x)) for x in pv.astype(object)]
    with open(output + ".fam", "w") as f:
        for i in range(n):
            f.write("\t".join([fam_cols["fam_id"][i], fam_cols["ind_id"][i], fam_cols["pat_id"][i], fam_cols["mat_id"][i],
                               sex[i], ph[i]]) + "\n")
    locus, alleles = dataset._rows["locus"], dataset._rows["alleles"]
    if varid is None:
        ids = [f"{l[0]}:{l[1]}:{a[0]}:{a[1]}" for l, a in zip(locus, alleles)]
    else:
        ids = _strings(varid, m, ".", "varid")
        bad = [v for v in ids if re.search(r"\s+", v)]
        if bad:
            raise TypeError("expr 'varid' has spaces in the following values:\n" + "".join(f"  {v}\n" for v in bad))
    cmv = np.zeros(m) if cm_position is None else np.nan_to_num(
        np.asarray(getattr(cm_position, "values", cm_position), dtype=np.float64), nan=0.0) * np.ones(m)
    with open(output + ".bim", "w") as f:
        for i in range(m):
            f.write(f"{locus[i][0]}\t{ids[i]}\t{cmv[i]}\t{locus[i][1]}\t{alleles[i][1]}\t{alleles[i][0]}\n")
    g = dataset.genotypes
    if isinstance(g, HostBedGenotypes):
        body = g.rows.numpy()[:, : (n + 3) // 4] if g.n_samples == n else None
    else:
        body = None
    if body is None:
        body = _bed_rows_of(dataset)
    with open(output + ".bed", "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(body).tobytes())


def _bed_rows_of(mt: MatrixTable) -> np.ndarray:
    """The .bed body of the dataset's CURRENT columns (ExportPlink bytes: hom-ref 11, het 10, hom-alt 00, missing 01)."""
    import torch

    from . import _lib

    g = mt.genotypes
    if isinstance(g, HostBedGenotypes):
        g = g.to_device()
    n_cols = mt.count_cols()
    if n_cols == g.n_samples and np.array_equal(mt.col_index, np.arange(n_cols)):
        ctx = _lib.context(g.device.index)
        stride = (n_cols + 3) // 4
        out = torch.empty((g.n_variants, stride), dtype=torch.uint8, device=g.device)
        with torch.cuda.device(g.device):
            # on torch's current stream: `g.data` was packed there, and .cpu() below synchronises with it
            ctx.check(ctx.lib.lrr_unpack_bed(ctx.handle, g.data.data_ptr(), g.stride, g.n_variants, g.n_samples,
                                             out.data_ptr(), stride, torch.cuda.current_stream(g.device).cuda_stream))
        return out.cpu().numpy()
    # filtered columns: re-encode the kept samples on the host (export is not on the hot path)
    dos = g.to_dosage()[:, mt.col_index]
    code = np.where(dos == 0, 3, np.where(dos == 1, 2, np.where(dos == 2, 0, 1))).astype(np.uint8)
    pad = (-code.shape[1]) % 4
    if pad:
        code = np.concatenate([code, np.zeros((code.shape[0], pad), dtype=np.uint8)], axis=1)
    code = code.reshape(code.shape[0], -1, 4)
    return (code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)).astype(np.uint8)


# ---- BGEN (v1.2, layout 2) ---------------------------------------------------------------------
def read_bgen(source, sample_file=None, want_probabilities=False):
    """Decode a BGEN file (path or bytes) on the host.  Returns a dict: `q` uint16 [M, N] (the dosage in units of 1/255:
    d1 + 2 d2 of the stored 8-bit probabilities, 0xFFFF = missing), `contig`, `position`, `alleles`, `rsid`, `varid`
    (file order), `samples`, and with `want_probabilities` the stored bytes `d0`, `d1` uint8 [M, N].

    Follows StagedBGENReader.scala: only what Hail reads is accepted (layout 2, two alleles, ploidy 2 for every sample,
    unphased, 8 bits per probability; zlib or no compression -- zstd needs a codec this image does not ship), with its
    fatal messages."""
    import struct
    import zlib
    b = source if isinstance(source, (bytes, bytearray, memoryview)) else open(source, "rb").read()
    b = bytes(b)
    if len(b) < 24:
        raise FatalError("BGEN file is too short to hold a header")
    offset, = struct.unpack_from("<I", b, 0)
    header_len, M, N = struct.unpack_from("<III", b, 4)
    magic = b[16:20]
    if magic not in (b"bgen", b"\0\0\0\0"):
        raise FatalError(f"expected magic number 'bgen' or 0000, found {magic!r}")
    flags, = struct.unpack_from("<I", b, 4 + header_len - 4)
    compression, layout, has_ids = flags & 3, (flags >> 2) & 15, (flags >> 31) & 1
    if layout != 2:
        raise FatalError(f"Hail only supports BGEN version 1.2 (layout 2), found layout {layout}")
    if compression not in (0, 1):
        raise NotImplementedError("read_bgen: zstd-compressed BGEN needs a zstd codec (only zlib / uncompressed here)")
    samples = None
    if has_ids:
        p = 4 + header_len
        _, n_ids = struct.unpack_from("<II", b, p)
        if n_ids != N:
            raise FatalError(f"BGEN file is malformed -- number of sample IDs in header does not equal number in file: {N}, {n_ids}")
        p += 8
        samples = []
        for _ in range(N):
            ln, = struct.unpack_from("<H", b, p)
            samples.append(b[p + 2:p + 2 + ln].decode())
            p += 2 + ln
    if sample_file is not None:   # impex.py:1103; LoadBgen.readSampleFile: two header lines, the id is column 1
        with open(sample_file) as f:
            lines = [ln.rstrip("\r\n") for ln in f if ln.strip()]
        ids = [re.split(r"\s+", ln)[0] for ln in lines[2:]]
        if len(ids) != N:
            raise FatalError(f"BGEN file and sample file have different numbers of samples: {N} vs {len(ids)}")
        samples = ids
    if samples is None:
        samples = [f"sample_{i}" for i in range(N)]   # (Hail: "sample_0", ... when the file carries no identifiers)
    q = np.empty((M, N), dtype=np.uint16)
    d0a = np.empty((M, N), dtype=np.uint8) if want_probabilities else None
    d1a = np.empty((M, N), dtype=np.uint8) if want_probabilities else None
    contig, position, alleles, rsid, varid = [], [], [], [], []
    blocks = []   # (offset of the genotype data, its size) per variant: the identifying data is walked serially ...
    p = 4 + offset
    for v in range(M):
        fields = []
        for _ in range(3):   # variant id, rsid, chromosome
            ln, = struct.unpack_from("<H", b, p)
            fields.append(b[p + 2:p + 2 + ln].decode())
            p += 2 + ln
        pos, n_alleles = struct.unpack_from("<IH", b, p)
        p += 6
        if n_alleles != 2:
            raise FatalError(f"Only biallelic variants supported, found variant with {n_alleles} alleles: {fields[2]}:{pos}")
        al = []
        for _ in range(n_alleles):
            ln, = struct.unpack_from("<I", b, p)
            al.append(b[p + 4:p + 4 + ln].decode())
            p += 4 + ln
        size, = struct.unpack_from("<I", b, p)
        p += 4
        blocks.append((p, size))
        p += size
        varid.append(fields[0])
        rsid.append(fields[1])
        contig.append(fields[2])
        position.append(int(pos))
        alleles.append(tuple(al))

    def decode(v):   # ... the probability blocks are decoded by a few threads (zlib and numpy release the GIL)
        p, size = blocks[v]
        where = f"{contig[v]}:{position[v]}"
        if compression:
            raw_len, = struct.unpack_from("<I", b, p)
            data = zlib.decompress(b[p + 4:p + size])
            if len(data) != raw_len:
                raise FatalError(f"BGEN block of {where} decompresses to {len(data)} bytes, header says {raw_len}")
        else:
            data = b[p:p + size]
        n_row, n_alleles2, min_ploidy, max_ploidy = struct.unpack_from("<IHBB", data, 0)
        if n_row != N:
            raise FatalError(f"Row nSamples is not equal to header nSamples: {n_row}, {N}")
        if n_alleles2 != 2:
            raise FatalError("Value for 'nAlleles' in genotype probability data storage is not equal to value in variant "
                             f"identifying data. Expected 2 but found {n_alleles2} at {where}.")
        if min_ploidy != 2 or max_ploidy != 2:
            raise FatalError(f"Hail only supports diploid genotypes. Found min ploidy '{min_ploidy}' and max ploidy '{max_ploidy}'.")
        ploidy = np.frombuffer(data, dtype=np.uint8, count=N, offset=8)
        bad = (ploidy & 0x3F) != 2
        if bad.any():
            raise FatalError(f"Ploidy value must equal to 2. Found {int(ploidy[bad][0])}.")
        phase, bits = data[8 + N], data[9 + N]
        if phase not in (0, 1):
            raise FatalError(f"Phase value must be 0 or 1. Found {phase}.")
        if phase == 1:
            raise FatalError("Hail does not support phased genotypes in 'import_bgen'.")
        if bits < 1 or bits > 32:
            raise FatalError(f"nBits value must be between 1 and 32 inclusive. Found {bits}.")
        if bits != 8:
            raise FatalError(f"Hail only supports 8-bit probabilities, found {bits}.")
        if len(data) != 2 * N + N + 10:
            raise FatalError(f"Number of uncompressed bytes '{len(data)}' does not match the expected size '{2 * N}'.")
        pr = np.frombuffer(data, dtype=np.uint8, count=2 * N, offset=10 + N).reshape(N, 2)
        d0, d1 = pr[:, 0].astype(np.int32), pr[:, 1].astype(np.int32)
        qv = (d1 + 2 * (255 - d0 - d1)).astype(np.uint16)      # (d1 + (d2 << 1)) / 255.0, StagedBGENReader.scala:350-353
        qv[(ploidy & 0x80) != 0] = CompactDosage.MISSING
        q[v] = qv
        if want_probabilities:
            d0a[v], d1a[v] = pr[:, 0], pr[:, 1]

    if M * N >= (1 << 22):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(decode, range(M)))   # (list: re-raises the first fatal condition)
    else:
        for v in range(M):
            decode(v)
    out = {"q": q, "contig": contig, "position": np.array(position, dtype=np.int64), "alleles": alleles, "rsid": rsid,
           "varid": varid, "samples": samples}
    if want_probabilities:
        out["d0"], out["d1"] = d0a, d1a
    return out


def import_bgen(path, entry_fields=("dosage",), sample_file=None, index_file_map=None, n_partitions=None, block_size=None,
                variants=None, *, reference_genome="default", contig_recoding=None, skip_invalid_loci=False,
                device=0) -> MatrixTable:
    """`hl.import_bgen` (impex.py:1100) for the regression path: entry field `dosage` as the compact uint16 store
    (`CompactDosage`, 2 bytes per entry on the device; `mt.dosage` is the float64 entry expression the reference yields,
    exact multiples of 1/255), row fields `locus`, `alleles`, `rsid`, `varid` (key locus, alleles; rows sorted by key),
    column key `s`.  `GT` / `GP` entry fields are not materialised on the device -- `read_bgen(...,
    want_probabilities=True)` returns the stored probabilities on the host.  `index_file_map`, `n_partitions`,
    `block_size` are accepted and inert (no index is needed to read the whole file); `variants` is not supported."""
    entry_fields = list(entry_fields)
    for f in entry_fields:
        if f not in ("GT", "GP", "dosage"):
            raise FatalError(f"Invalid entry field '{f}'. Expected one of 'GT', 'GP', 'dosage'.")   # impex.py import_bgen
    if "dosage" not in entry_fields:
        raise NotImplementedError("import_bgen: the device store holds the 'dosage' entry field; ask for it "
                                  "(GT / GP: read_bgen(..., want_probabilities=True))")
    if variants is not None:
        raise NotImplementedError("import_bgen: variants= (index-based filtering) is not supported")
    if reference_genome == "default":
        reference_genome = "GRCh37"
    if reference_genome not in (None, "GRCh37"):
        raise NotImplementedError("import_bgen: only reference_genome='GRCh37' or None")
    d = read_bgen(path, sample_file)
    recode = contig_recoding or {}
    rank = {c: i for i, c in enumerate(_GRCH37_CONTIGS)} if reference_genome is not None else None
    keep, contig = [], []
    for i, (c, pos) in enumerate(zip(d["contig"], d["position"])):
        c = recode.get(c, c)
        valid = rank is None or (c in rank and pos >= 1)
        if not valid and not skip_invalid_loci:
            raise FatalError(f"Invalid locus '{c}:{pos}' found. Contig '{c}' is not in the reference genome '{reference_genome}'.")
        if valid:
            keep.append(i)
            contig.append(c)
    keep = np.array(keep, dtype=np.int64)
    pos = d["position"][keep]
    alleles = [d["alleles"][i] for i in keep]
    keys = [((rank[c] if rank is not None else c), int(p_), a) for c, p_, a in zip(contig, pos, alleles)]
    order = np.array(sorted(range(len(keys)), key=keys.__getitem__), dtype=np.int64)
    rows_sel = keep[order]
    store = CompactDosage.from_q(np.ascontiguousarray(d["q"][rows_sel]), 1.0 / 255.0, device)
    take = lambda seq: [seq[i] for i in rows_sel]
    rows = {
        "locus": np.array([(contig[i], int(pos[i])) for i in order] + [None], dtype=object)[:-1],
        "alleles": np.array([alleles[i] for i in order] + [None], dtype=object)[:-1],
        "rsid": np.array(take(d["rsid"]), dtype=object),
        "varid": np.array(take(d["varid"]), dtype=object),
    }
    return MatrixTable(store, rows=rows, cols={"s": np.array(d["samples"], dtype=object)}, row_key=("locus", "alleles"),
                       col_key=("s",))
