"""PLINK ingest / export for the B200 path: .bed/.bim/.fam <-> MatrixTable with packed GT.

Reference behaviour followed:
  * `hl.import_plink` / `hl.import_fam`    hail/python/hail/methods/impex.py:2505 (defaults: a2_reference=True, missing='NA',
                                           delimiter='\\\\s+', quant_pheno=False); driver checks and messages
                                           hail/hail/src/is/hail/io/plink/LoadPlink.scala:40-82 (bim), :102-186 (fam),
                                           :225-251 (magic, SNP-major, file size), :475-481 + :525 (entry decode);
                                           rows are SORTED by (locus, alleles) (:79-81), each keeping its .bed row index
  * `hl.export_plink`                      impex.py:324-470 (defaults, white-space check),
                                           hail/hail/src/is/hail/expr/ir/MatrixWriter.scala:2110-2285 (bytes written)
The genotype bytes never pass through Python objects: they are read into page-locked memory (`resident=False`: the
out-of-core form streamed by lrr_stream_*) or packed into the device store (`resident=True`).
"""
from __future__ import annotations

import os
import re

import numpy as np

from .genotypes import BED_MAGIC, HostBedGenotypes, PackedGenotypes
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
