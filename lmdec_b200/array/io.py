"""PLINK input and operator persistence (reference lmdec/array/io.py:16-336), without pandas_plink / zarr / dask.

``load_plink_array`` keeps the reference's signature and orientation rules and returns a device-resident
``GenotypeStore``: the ``.bed`` payload is already 2-bit, so it is recoded on the GPU (``lmdec_recode_bed`` +
``lmdec_store_transpose``) with no float round trip.

``save_array`` / ``load_array_from_disk`` keep the reference's TREE layout (``chained/``, ``stacked/``, ``<i>_array``,
``<i>_sparse`` with ``coords / data / shape``), so an operator saved here opens with the reference's loader and vice
versa -- for the leaf formats both sides have: ``npy`` and ``txt`` (zarr, the reference's default, is not in this
image).  ``dense_file_format='packed'`` (the default for a ``GenotypeOperator``) replaces the int8 genotype leaf and its
sparse mask by ONE 2-bit leaf ``<i>_packed/`` (16x smaller than int8; holds the tile store, the shape, the integer code
counts and the imputation vector): that leaf is this build's own and the reference cannot read it.

.bed code map -- PARITY UNPINNED (the reference delegates to pandas_plink and has no test for it):
2-bit field (hi lo): 00 -> 0, 01 -> missing, 10 -> 1, 11 -> 2 (count of the .bim's second allele, what
``pandas_plink.read_plink`` yields); ``count='a1'`` flips 0 <-> 2 (what ``read_plink1_bin`` yields with its default
``ref='a1'``).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from ..device import F64, DeviceArray, require_cuda
from ..store import GenotypeOperator, GenotypeStore

BED_MAGIC = bytes([0x6C, 0x1B])


def _count_lines(path) -> int:
    n = 0
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            n += chunk.count(b"\n")
        if f.tell() > 0:
            f.seek(-1, os.SEEK_END)
            if f.read(1) != b"\n":
                n += 1
    return n


def read_bed(bed: Union[str, Path], n_samples: int, n_variants: int, count: str = "a0", device=None) -> GenotypeStore:
    """``.bed`` file -> GenotypeStore with rows = VARIANTS and columns = samples (the matrix ``read_plink`` returns).
    Checks the 3-byte header (0x6c 0x1b, then 0x01 = variant-major / 0x00 = sample-major) and the payload size."""
    if count not in ("a0", "a1"):
        raise ValueError("count must be 'a0' or 'a1'")
    device = device or require_cuda()
    raw = np.fromfile(str(bed), dtype=np.uint8)
    if raw.size < 3 or bytes(raw[:2]) != BED_MAGIC:
        raise ValueError(f"{bed}: not a PLINK 1 .bed file (bad magic bytes)")
    mode = int(raw[2])
    if mode not in (0, 1):
        raise ValueError(f"{bed}: unknown .bed mode byte {mode}")
    major, minor = (n_variants, n_samples) if mode == 1 else (n_samples, n_variants)
    payload = raw[3:]
    if payload.size != major * ((minor + 3) // 4):
        raise ValueError(f"{bed}: payload of {payload.size} bytes does not match {n_variants} variants x "
                         f"{n_samples} samples from the .bim / .fam files")
    t = torch.from_numpy(payload).to(device)
    if count == "a1":
        # 00 <-> 11, 01 and 10 unchanged: invert both bits of every field whose two bits are equal
        same = ~((t >> 1) ^ t) & 0x55
        t = t ^ (same | (same << 1))
    if mode == 1:
        return GenotypeStore.from_bed_bytes(t, n_samples=n_samples, n_variants=n_variants, device=device,
                                            variants_as_rows=True)
    # sample-major payload: the same recode with the roles of the axes swapped
    return GenotypeStore.from_bed_bytes(t, n_samples=n_variants, n_variants=n_samples, device=device,
                                        variants_as_rows=False)


def load_plink_array(path_to_plink_files: Optional[Union[str, Path]] = None, bed: Optional[Union[str, Path]] = None,
                     bim: Optional[Union[str, Path]] = None, fam: Optional[Union[str, Path]] = None,
                     transpose: bool = False, device=None) -> GenotypeStore:
    """Genotypes of a PLINK 1 fileset as a device-resident 2-bit store (reference lmdec/array/io.py:16-73).

    Same argument rules as the reference: either ``path_to_plink_files`` ('/path/to/data' for data.bed/.bim/.fam) or all
    of ``bed``, ``bim``, ``fam``; anything else raises the reference's ValueError.  Orientation as in the reference: the
    prefix form returns what ``pandas_plink.read_plink`` returns, VARIANTS x samples with the a0 count; the three-file
    form returns what ``read_plink1_bin(...).data`` returns, SAMPLES x variants with the a1 count; ``transpose=True``
    swaps the axes.  Feed the result to ``make_snp_array`` (column moments, mean imputation) or, when nothing is
    missing, straight to ``PowerMethod.svd``."""
    if path_to_plink_files is not None and not all([bed, bim, fam]):
        prefix = str(path_to_plink_files)
        for ext in (".bed", ".bim", ".fam"):
            if prefix.endswith(ext):
                prefix = prefix[:-4]
        bed_f, bim_f, fam_f = prefix + ".bed", prefix + ".bim", prefix + ".fam"
        count, samples_as_rows = "a0", False
    elif all(p is not None for p in [bed, bim, fam]) and not path_to_plink_files:
        bed_f, bim_f, fam_f = str(bed), str(bim), str(fam)
        count, samples_as_rows = "a1", True
    else:
        raise ValueError("Uninterpretable input."
                         " Please specify array, xor path_to_files, xor (bed and bim and fim)")
    n_variants, n_samples = _count_lines(bim_f), _count_lines(fam_f)
    store = read_bed(bed_f, n_samples, n_variants, count=count, device=device)      # variants x samples
    if samples_as_rows != bool(transpose):
        store = store.transposed()
    return store


# ------------------------------------------------------------------------------------------------- persistence
def _save_leaf(array: np.ndarray, file: str, file_format: str):
    if file_format == "npy":
        np.save(file if file.endswith(".npy") else file + ".npy", array)
    elif file_format == "txt":
        np.savetxt(file if file.endswith(".txt") else file + ".txt", array)
    else:
        raise ValueError(f"file_format, {file_format}, not implemented (this image has no zarr: use 'npy' or 'txt')")


def _save_sparse(coords: np.ndarray, data: np.ndarray, shape, path: str, file_format: str):
    if file_format not in ("npy", "txt"):
        raise ValueError(f"expected file_format in ['npy', 'txt'], but got {file_format}")
    os.mkdir(path)
    _save_leaf(coords, path + "/coords", file_format)
    _save_leaf(data, path + "/data", file_format)
    _save_leaf(np.asarray(shape), path + "/shape", file_format)


def _save_packed(op: GenotypeOperator, path: str):
    """<i>_packed/: the sample-major tile store (the SNP-major one is re-derived by lmdec_store_transpose on load),
    the exact integer code counts (moments are functions of them) and the imputation vector."""
    os.mkdir(path)
    st = op.store
    np.save(path + "/sample_major.npy", st.s_store.cpu().numpy())
    np.save(path + "/code_counts.npy", st.code_counts().cpu().numpy())
    if op.has_missing:
        np.save(path + "/impute.npy", op.impute_mean.cpu().numpy())
    with open(path + "/meta.json", "w") as f:
        json.dump({"format": "lmdec_b200 tile store v1: 2-bit dosage codes (3 = missing), 128 x 256 tiles",
                   "n": st.n, "p": st.p, "rows": op.rows, "planes": op.planes}, f)


def save_array(array, file: Union[str, Path], dense_file_format: Optional[str] = None,
               sparse_file_format: str = "txt") -> None:
    """Save a genotype operator / Stacked / Chained / dense array as the reference's tree (io.py:177-273).

    A ``GenotypeOperator`` is written under its container spelling  chained[ stacked[ stacked[G, M], -mean ], 1/std ]
    (levels without a vector are dropped, as ``make_snp_array`` does), so mean, 1/std and the imputation values are
    persisted next to the genotypes and a load does not refit anything.  ``dense_file_format``: 'packed' (default for
    genotype operators: one 2-bit leaf for G and M), or 'npy' / 'txt' (the reference's own leaves: int8 G plus the
    sparse mask as coords / data / shape -- readable by the reference's ``load_array_from_disk``)."""
    from .chained import ChainedArray, _Diag
    from .operators import DenseOperator, SparseOperator
    from .stacked import StackedArray, _Vector
    file = str(file)
    if os.path.exists(file):
        raise ValueError(f"file, {file} already exists")
    if isinstance(array, GenotypeStore):
        array = GenotypeOperator(array, None, None, impute=False) if not array.has_missing else \
            GenotypeOperator(array, None, None)
    if isinstance(array, GenotypeOperator):
        fmt = dense_file_format or "packed"
        if array.store.group is not None or array.store.snp_group is not None:
            raise NotImplementedError("save_array writes one rank's operator; gather or save per rank")
        vec_fmt = "npy" if fmt == "packed" else fmt
        if fmt == "packed":
            node = ("packed", array)
        else:
            g = array.store.to_dense(rows=array.rows)
            if array.has_missing:
                nan = g < 0
                rows, cols = torch.nonzero(nan, as_tuple=True)
                mask = (torch.stack([rows, cols]).cpu().numpy(), array.impute_mean[cols].cpu().numpy(),
                        (array.rows, array.store.p))
                node = ("stacked", [("array", torch.where(nan, torch.zeros_like(g), g).cpu().numpy(), fmt),
                                    ("sparse", mask)])
            else:
                node = ("array", g.cpu().numpy(), fmt)
        if array.mean is not None:
            node = ("stacked", [node, ("array", (-array.mean).cpu().numpy(), vec_fmt)])
        if array.inv_std is not None:
            node = ("chained", [node, ("array", array.inv_std.cpu().numpy(), vec_fmt)])

        def write(nd, path, top):
            kind = nd[0]
            if top:
                os.mkdir(path)
                path = path + "/" + kind
            if kind in ("stacked", "chained"):
                os.mkdir(path)
                for i, child in enumerate(nd[1]):
                    write(child, f"{path}/{i}_{child[0]}", False)
            elif kind == "array":
                _save_leaf(nd[1], path, nd[2])
            elif kind == "sparse":
                _save_sparse(*nd[1], path, sparse_file_format)
            else:
                _save_packed(nd[1], path)

        write(node, file, True)
        return
    fmt = dense_file_format or "npy"
    if isinstance(array, (StackedArray, ChainedArray)):
        os.mkdir(file)
        if not (file.endswith("stacked") or file.endswith("chained")):
            file = file + ("/stacked" if isinstance(array, StackedArray) else "/chained")
            os.mkdir(file)
        for i, a in enumerate(array.arrays):
            if isinstance(a, StackedArray):
                save_array(a, f"{file}/{i}_stacked", fmt, sparse_file_format)
            elif isinstance(a, ChainedArray):
                save_array(a, f"{file}/{i}_chained", fmt, sparse_file_format)
            elif isinstance(a, GenotypeOperator):
                _save_packed(a, f"{file}/{i}_packed")
            elif isinstance(a, SparseOperator):
                coo = a._host.tocoo()
                _save_sparse(np.stack([coo.row, coo.col]), coo.data, coo.shape, f"{file}/{i}_sparse",
                             sparse_file_format)
            elif isinstance(a, (_Vector, _Diag)):
                _save_leaf(a.v.cpu().numpy().reshape(a.shape), f"{file}/{i}_array", fmt)
            elif isinstance(a, DenseOperator):
                _save_leaf(a.a.cpu().numpy(), f"{file}/{i}_array", fmt)
            else:
                raise ValueError(f"cannot save a member of type {type(a)}")
        return
    if isinstance(array, SparseOperator):
        coo = array._host.tocoo()
        if not file.endswith("sparse"):
            os.mkdir(file)
            file = file + "/sparse"
        _save_sparse(np.stack([coo.row, coo.col]), coo.data, coo.shape, file, sparse_file_format)
        return
    dense = array.compute() if hasattr(array, "compute") else np.asarray(array)
    if not file.endswith("array"):
        os.mkdir(file)
        file = file + "/array"
    _save_leaf(dense, file, fmt)


def sort_files_by_int(x) -> List[str]:
    return list(sorted(x, key=lambda s: int(s.split("_")[0])))


def walk_one_level(top, ignore_hidden: bool = True) -> Tuple[List[str], List[str]]:
    for (_, dirs, files) in os.walk(str(top)):
        if ignore_hidden:
            files = [f for f in files if not f.startswith(".")]
        return list(dirs), list(files)
    return [], []


def _load_dense(file: Path):
    if file.suffix == ".npy":
        return np.load(str(file))
    if file.suffix == ".txt":
        return np.loadtxt(str(file))
    raise ValueError(f"{file} does not reference an load-able array")


def _load_sparse(path: Path):
    import scipy.sparse as sp
    _, files = walk_one_level(path)
    pick = lambda stem: Path(path, [f for f in files if f.startswith(stem)][0])     # noqa: E731
    shape = tuple(int(i) for i in np.atleast_1d(_load_dense(pick("shape"))))
    coords = np.asarray(_load_dense(pick("coords"))).reshape(len(shape), -1).astype(np.int64)
    data = np.atleast_1d(_load_dense(pick("data")))
    return sp.coo_matrix((data, (coords[0], coords[1])), shape=shape)


def _load_packed(path: Path, device=None) -> GenotypeOperator:
    with open(Path(path, "meta.json")) as f:
        meta = json.load(f)
    store = GenotypeStore(meta["n"], meta["p"], device=device)
    s = torch.from_numpy(np.load(str(Path(path, "sample_major.npy"))))
    if s.numel() != store.s_store.numel():
        raise ValueError("stored tile store does not match its recorded shape")
    store.s_store.copy_(s)
    store.derive_snp_major()
    counts = torch.from_numpy(np.load(str(Path(path, "code_counts.npy")))).to(store.device)
    if not torch.equal(store.code_counts(), counts):
        raise ValueError(f"{path}: code counts of the loaded store differ from the saved ones (corrupt file)")
    op = GenotypeOperator(store, None, None, planes=meta.get("planes", 3), rows=meta.get("rows"))
    imp = Path(path, "impute.npy")
    if imp.exists():
        op.impute_mean = torch.from_numpy(np.load(str(imp))).to(store.device)
    return op


def load_array_from_disk(path_to_array: Union[str, Path], device=None, recognise: bool = True):
    """Inverse of ``save_array`` (io.py:155-174): rebuilds the Stacked / Chained tree; a tree that spells the genotype
    operator (with a packed leaf, or with the reference's int8 + sparse-mask leaves) comes back as a
    ``GenotypeOperator`` with the SAVED mean / 1/std / imputation vector (``recognise=False`` keeps the containers)."""
    from .chained import ChainedArray
    from .stacked import StackedArray
    path = Path(path_to_array)
    name = str(path)
    if not os.path.isdir(path):
        return _load_dense(path)
    if name.endswith("sparse"):
        return _load_sparse(path)
    if name.endswith("packed"):
        return _load_packed(path, device)
    if name.endswith("stacked") or name.endswith("chained"):
        dirs, files = walk_one_level(path)
        members = [load_array_from_disk(Path(path, s), device=device, recognise=False)
                   for s in sort_files_by_int(dirs + files)]
        out = (StackedArray if name.endswith("stacked") else ChainedArray)(members, device=device)
        return _fold(out) if recognise else out
    dirs, files = walk_one_level(path)
    if len(dirs) == 1:
        return load_array_from_disk(Path(path, dirs[0]), device=device, recognise=recognise)
    if len(files) == 1 and not dirs:
        return _load_dense(Path(path, files[0]))
    raise ValueError(f"{path} is not an array tree written by save_array")


def _fold(tree):
    """container tree -> GenotypeOperator when it is the make_snp_array pattern (packed or int8 + mask leaves)"""
    from .chained import ChainedArray, _Diag
    from .recognize import genotype_operator_from_containers
    from .stacked import StackedArray, _Vector
    node, inv_std, mean = tree, None, None
    if isinstance(node, ChainedArray) and len(node.arrays) == 2 and isinstance(node.arrays[1], _Diag):
        node, inv_std = node.arrays[0], node.arrays[1].v
    if isinstance(node, StackedArray) and len(node.arrays) == 2 and isinstance(node.arrays[1], _Vector):
        node, mean = node.arrays[0], -node.arrays[1].v
    if isinstance(node, GenotypeOperator):
        op = GenotypeOperator(node.store, mean, inv_std, planes=node.planes, rows=node.rows)
        if node.has_missing:
            op.impute_mean = node.impute_mean
        return op
    op = genotype_operator_from_containers(tree)
    return tree if op is None else op
