"""Phantom dump -> particle columns RESIDENT ON THE GPU.

The step before the scatter path.  Sarracen's `read_phantom`
(sarracen/readers/read_phantom.py:173-269, :286-450) reads every array block into a pandas
DataFrame; rendering it then costs a `.to_numpy()` and a host -> device copy per column and per
call.  Here the dump is memory-mapped, its Fortran record structure is indexed without touching
the particle data, and every array block is streamed from the page cache through page-locked
slots (one worker thread per column, `backend._stage_column`) straight into device tensors;
inactive particles (h <= 0) are removed, and the mass column is built, on the device.  The result
is a `DeviceFrame`: what `sarracen_b200.interpolate` needs of a data frame, without pandas.

File format (https://phantomsph.readthedocs.io/en/latest/dumpfile.html): Fortran unformatted
records -- a 4-byte length tag before and after each record.  Capture-pattern record
(i1 = 60769, r1 = 60878, i2 = 60878, iversion, i3 = 690706 in the file's default int / real kinds),
100-character file identifier, global header (for each of the 8 kinds default-int, int8, int16,
int32, int64, default-real, real4, real8: count; 16-character tags; values), number of array
blocks, per MPI block and array block (n as int64, 8 counts as int32), then per array a tag
record and a data record.  Block 2 of every MPI block holds the sink particles.  Like the
reference, the record tags are only checked for consistency (leading == trailing); sizes come
from the structure.  Either byte order is read.
"""
import mmap
import os
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from .frame import DeviceFrame, ParticleFrame

_KINDS = ("i_def", "i1", "i2", "i4", "i8", "r_def", "r4", "r8")


class _Walker:
    """Cursor over the Fortran records of a memory-mapped dump."""

    def __init__(self, buf):
        self.buf, self.pos, self.size = buf, 0, len(buf)

    def record(self, nbytes: int) -> Tuple[int, int]:
        """(offset, nbytes) of the payload of the next record, which must hold `nbytes` bytes."""
        end = self.pos + 8 + nbytes
        if nbytes < 0 or end > self.size:
            raise AssertionError("unexpected end of file: is this a Phantom data file?")
        if self.buf[self.pos:self.pos + 4] != self.buf[end - 4:end]:
            raise AssertionError("Fortran tags mismatch.")
        off = self.pos + 4
        self.pos = end
        return off, nbytes

    def array(self, nbytes: int, dtype) -> np.ndarray:
        off, _ = self.record(nbytes)
        return np.frombuffer(self.buf, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=off)

    def text(self, nbytes: int) -> str:
        off, _ = self.record(nbytes)
        return bytes(self.buf[off:off + nbytes]).decode("ascii")


def _capture_pattern(w: _Walker):
    """Default int / real kinds and byte order from the first record (read_phantom.py:27-100)."""
    for isz, rsz in ((4, 8), (4, 4), (8, 8), (8, 4)):
        for order in ("<", ">"):
            it, rt = np.dtype(f"{order}i{isz}"), np.dtype(f"{order}f{rsz}")
            if 4 + 4 * isz + rsz + 4 > w.size:
                continue
            i1 = np.frombuffer(w.buf, it, 1, 4)[0]
            r1 = np.frombuffer(w.buf, rt, 1, 4 + isz)[0]
            i2 = np.frombuffer(w.buf, it, 1, 4 + isz + rsz)[0]
            if i1 == 60769 and i2 == 60878 and r1 == float(i2):
                iversion = np.frombuffer(w.buf, it, 1, 4 + 2 * isz + rsz)[0]
                i3 = np.frombuffer(w.buf, it, 1, 4 + 3 * isz + rsz)[0]
                if i3 != 690706:
                    raise AssertionError("Capture pattern error. i3 mismatch. Is this a Phantom data file?")
                w.record(4 * isz + rsz)
                return it, rt, int(iversion), order
    raise AssertionError("Could not determine default int or float precision (i1, r1, i2 mismatch). "
                         "Is this a Phantom data file?")


def _kind_dtypes(it, rt, order):
    return [it, np.dtype("i1"), np.dtype(f"{order}i2"), np.dtype(f"{order}i4"), np.dtype(f"{order}i8"),
            rt, np.dtype(f"{order}f4"), np.dtype(f"{order}f8")]


def _native(value):
    """Header scalars as native-endian NumPy scalars (what the reference hands out)."""
    arr = np.asarray(value)
    return arr.astype(arr.dtype.newbyteorder("="))[()]


def _global_header(w: _Walker, dtypes, order) -> dict:
    keys, vals = [], []
    for dt in dtypes:
        nvars = int(w.array(4, np.dtype(f"{order}i4"))[0])
        if nvars > 0:
            tags = w.text(16 * nvars)
            keys += [tags[i:i + 16].strip() for i in range(0, 16 * nvars, 16)]
            vals += [_native(v) for v in w.array(dt.itemsize * nvars, dt)]
    seen: Dict[str, int] = {}
    out = {}
    for k, v in zip(keys, vals):   # duplicates become key_2, key_3, ... (read_phantom.py:113-124)
        seen[k] = seen.get(k, 0) + 1
        out[k if seen[k] == 1 else f"{k}_{seen[k]}"] = v
    return out


def _index_arrays(w: _Walker, dtypes, order, mpi_blocks: int):
    """Where every particle array lives in the file: ({tag: [(offset, dtype, n), ...]}, same for sinks);
    one piece per MPI block.  No particle data is touched."""
    nblocks = int(w.array(4, np.dtype(f"{order}i4"))[0]) // max(1, mpi_blocks)
    gas: Dict[str, list] = {}
    sinks: Dict[str, list] = {}
    for _ in range(max(1, mpi_blocks)):
        counts = []
        for _b in range(nblocks):
            off, _ = w.record(40)
            n = int(np.frombuffer(w.buf, np.dtype(f"{order}i8"), 1, off)[0])
            nums = np.frombuffer(w.buf, np.dtype(f"{order}i4"), 8, off + 8).astype(np.int64)
            counts.append((n, nums))
        for b, (n, nums) in enumerate(counts):
            target = sinks if b == 1 else gas     # block 2 is the sink particles (read_phantom.py:246-256)
            local = set()
            for kind, dt in enumerate(dtypes):
                for _a in range(int(nums[kind])):
                    tag = w.text(16).strip()
                    name, count = tag, 1
                    while name in local:          # repeated tags inside a block: tag_2, tag_3, ...
                        count += 1
                        name = f"{tag}_{count}"
                    local.add(name)
                    off, nbytes = w.record(dt.itemsize * n)
                    target.setdefault(name, []).append((off, dt, n))
    return gas, sinks


def _host_column(buf, pieces) -> np.ndarray:
    """A (small) array assembled on the host, native byte order."""
    parts = [np.frombuffer(buf, dt, n, off) for off, dt, n in pieces]
    arr = parts[0] if len(parts) == 1 else np.concatenate(parts)
    return arr.astype(arr.dtype.newbyteorder("="))


def _device_columns(buf, index: Dict[str, list], device, columns=None):
    """Stream the indexed arrays into device tensors.  CUDA: page cache -> page-locked slots ->
    device, one worker thread per column, all columns in flight together; byte order fixed on the
    device.  device='cpu' (tests, tools): plain host tensors."""
    import torch
    from . import backend
    dev = torch.device(device)
    names = [k for k in index if columns is None or k in columns]
    out = {}
    if dev.type != "cuda":
        for k in names:
            out[k] = torch.from_numpy(np.ascontiguousarray(_host_column(buf, index[k])))
        return out
    from concurrent.futures import ThreadPoolExecutor
    if backend._stage_pool is None:
        backend._stage_pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix="sphb-stage")
    copy_stream = backend._copy_streams.get(dev.index)
    if copy_stream is None:
        copy_stream = backend._copy_streams[dev.index] = torch.cuda.Stream(dev)
    jobs, swapped = [], []
    with torch.cuda.device(dev):
        for j, k in enumerate(names):
            pieces = index[k]
            dt = pieces[0][1]
            native = np.dtype(dt.str.replace(">", "<").replace("|", "<") if dt.itemsize > 1 else dt.str)
            tdt = torch.from_numpy(np.zeros(1, native)).dtype
            total = sum(n for _o, _d, n in pieces)
            dst = torch.empty(total, dtype=tdt, device=dev)
            out[k] = dst
            if dt.byteorder == ">" and dt.itemsize > 1:
                swapped.append(k)
            lo = 0
            for off, _d, n in pieces:
                # the bytes as they are in the file, viewed in the native kind (swapped on the device below)
                src = backend._from_numpy(np.frombuffer(buf, native, n, off))   # read-only view of the map
                jobs.append(backend._stage_pool.submit(backend._stage_column, j % 8, src, dst[lo:lo + n],
                                                       copy_stream))
                lo += n
        for job in jobs:
            job.result()
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        for k in swapped:
            t = out[k]
            out[k] = t.view(torch.uint8).view(-1, t.element_size()).flip(1).contiguous().view(t.dtype).view(-1)
    return out


def read_phantom(filename: str, separate_types: Optional[str] = "sinks", ignore_inactive: bool = True,
                 device=None, columns=None) -> Union[DeviceFrame, List]:
    """Read a Phantom dump into GPU-resident columns.

    Same arguments and return structure as `sarracen.read_phantom` (read_phantom.py:286-450):
    `separate_types` None / 'sinks' / 'all', `ignore_inactive` drops particles with h <= 0; the
    global header goes to `.params`, with 'mass' set from 'massoftype' (or a 'mass' column built
    for mixed particle types / APR levels).  SPH particles come back as `DeviceFrame`s (columns on
    `device`, default the current CUDA device), sink particles -- a handful -- as a host
    `ParticleFrame`.  `columns` restricts the particle arrays that are loaded."""
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("read_phantom needs a CUDA device (or an explicit device='cpu')")
        device = torch.device("cuda", torch.cuda.current_device())
    with open(filename, "rb") as fp:
        size = os.fstat(fp.fileno()).st_size
        if size == 0:
            raise AssertionError("empty file: is this a Phantom data file?")
        buf = mmap.mmap(fp.fileno(), 0, access=mmap.ACCESS_READ)
    try:
        w = _Walker(buf)
        it, rt, iversion, order = _capture_pattern(w)
        identifier = w.text(100).strip()
        dtypes = _kind_dtypes(it, rt, order)
        params = _global_header(w, dtypes, order)
        params["file_identifier"] = identifier
        params["iversion"] = iversion
        params["def_int_dtype"] = np.dtype(it.str.lstrip("<>=")).type
        params["def_real_dtype"] = np.dtype(rt.str.lstrip("<>=")).type
        mpi_blocks = int(params["nblocks"]) if "nblocks" in params else 1
        gas_index, sink_index = _index_arrays(w, dtypes, order, mpi_blocks)
        cols = _device_columns(buf, gas_index, device, columns)
        sinks = {k: _host_column(buf, v).copy() for k, v in sink_index.items()}
    finally:
        try:
            buf.close()
        except BufferError:   # NumPy views of the map are still alive; it is unmapped when they go
            pass

    if ignore_inactive and "h" in cols:
        keep = cols["h"] > 0
        if not bool(keep.all()):
            cols = {k: v[keep] for k, v in cols.items()}
    mixed = "itype" in cols and int(torch.unique(cols["itype"]).numel()) > 1
    if separate_types != "all" and mixed:
        # one frame with several species: a mass column (read_phantom.py:272-283)
        mass = torch.full_like(cols["itype"], float(params["massoftype"]), dtype=torch.float64)
        for itype in torch.unique(cols["itype"]).tolist():
            if itype > 1:
                mass[cols["itype"] == itype] = float(params[f"massoftype_{itype}"])
        cols["mass"] = mass
    elif "apr_level" in cols:
        cols["mass"] = float(params["massoftype"]) / torch.pow(
            torch.tensor(2.0, dtype=torch.float64, device=cols["apr_level"].device),
            cols["apr_level"].to(torch.float64) - 1.0)
    else:
        params["mass"] = params["massoftype"]

    frames: List = []
    if separate_types == "all" and mixed:
        for itype in sorted(torch.unique(cols["itype"]).tolist()):
            sel = cols["itype"] == itype
            key = "massoftype" if itype == 1 else f"massoftype_{itype}"
            frames.append(DeviceFrame({k: v[sel] for k, v in cols.items()}, {**params, "mass": params[key]}))
    elif separate_types in ("all", "sinks"):
        frames.append(DeviceFrame(cols, params))
    if separate_types in ("all", "sinks"):
        if sinks and len(next(iter(sinks.values()))) > 0:
            frames.append(ParticleFrame(sinks, params={k: v for k, v in params.items() if k != "mass"}))
        return frames[0] if len(frames) == 1 else frames
    # None: everything in one frame; the sink rows are appended (columns a sink lacks are NaN)
    if sinks and len(next(iter(sinks.values()))) > 0:
        n_s = len(next(iter(sinks.values())))
        merged = {}
        for k in list(cols) + [k for k in sinks if k not in cols]:
            a = cols.get(k)
            ref = a if a is not None else next(iter(cols.values()))
            if k in sinks:
                b = torch.from_numpy(np.ascontiguousarray(sinks[k])).to(ref.device)
            else:
                b = torch.full((n_s,), float("nan"), dtype=torch.float64, device=ref.device)
            if a is None:
                a = torch.full((ref.numel(),), float("nan"), dtype=torch.float64, device=ref.device)
            if a.dtype != b.dtype:
                a, b = a.to(torch.float64), b.to(torch.float64)
            merged[k] = torch.cat([a, b])
        cols = merged
    return DeviceFrame(cols, params)
