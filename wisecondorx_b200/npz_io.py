"""Reference `.npz` I/O on all host cores.

The reference writes its files with np.savez_compressed (main.py:33, newref_control.py:145,176,237): a zip archive
with one deflate stream per array, produced by a single thread -- at 15 kb / 500 samples the final reference holds
3 x (indexes 0.23 GB, distances 0.46 GB, null_ratios 0.15 GB) and zlib alone takes longer than every GPU stage
together.  `savez_compressed` here writes the SAME format (np.load reads it back, the reference reads it) but

  * compresses the arrays concurrently, and
  * splits a large array into blocks that are deflated independently and concatenated into one valid raw deflate
    stream (every block but the last ends on a full-flush boundary, the last one finishes the stream; the CRC-32 of
    the member is the chained crc32 of the blocks) -- the pigz construction, zlib releases the GIL while it works.

`load_samples` reads many sample files concurrently (inflate also releases the GIL)."""
from __future__ import annotations

import io
import os
import struct
import zlib
import zipfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

BLOCK = 8 << 20  # bytes deflated per task


def _npy_parts(arr):
    """The .npy serialisation of one array, exactly what np.savez stores in a member, as (header bytes, data
    buffer): C-contiguous numeric arrays are not copied, everything else goes through np.lib.format.write_array."""
    arr = np.asanyarray(arr)
    if arr.dtype.hasobject or not arr.flags.c_contiguous or arr.ndim == 0 or arr.size == 0:
        bio = io.BytesIO()
        np.lib.format.write_array(bio, arr, allow_pickle=True)
        return b"", bio.getbuffer()
    bio = io.BytesIO()
    np.lib.format.write_array_header_1_0(bio, np.lib.format.header_data_from_array_1_0(arr))
    return bio.getvalue(), memoryview(arr).cast("B")


def _member_blocks(header: bytes, data):
    """BLOCK-sized pieces of header + data; only the first piece (which carries the header) is a copy."""
    total = len(header) + len(data)
    if total <= BLOCK or not header:
        whole = (header + bytes(data)) if header else data
        return [whole[b * BLOCK:(b + 1) * BLOCK] for b in range(max(1, -(-len(whole) // BLOCK)))], total
    first = BLOCK - len(header)
    blocks = [header + bytes(data[:first])]
    for o in range(first, len(data), BLOCK):
        blocks.append(data[o:o + BLOCK])
    return blocks, total


def _store_ratio() -> float:
    try:
        return float(os.environ.get("WCX_NPZ_STORE_RATIO", "0.85"))
    except ValueError:
        return 0.85


def _deflate_block(args):
    """One independently deflated block.  A block whose first 64 KB do not shrink below WCX_NPZ_STORE_RATIO (0.85) at
    level 1 is emitted as stored deflate blocks (level 0): the float64 distances and null ratios of a reference file
    only compress to 0.90 / 0.96, and deflating them was the critical path of `newref` (34 core-seconds at 15 kb).  It
    is still an ordinary deflate stream for every reader.  Ratio >= 1 disables the test."""
    buf, last, level = args
    if level > 0 and len(buf) >= (1 << 16):
        ratio = _store_ratio()
        if ratio < 1.0 and len(zlib.compress(buf[:1 << 16], 1)) > ratio * (1 << 16):
            level = 0
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    out = c.compress(buf)
    out += c.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH)
    return out, zlib.crc32(buf), len(buf)


def savez_compressed(file, threads: int | None = None, level: int = 6, **arrays):
    """Drop-in for np.savez_compressed(file, **arrays) with parallel deflate; same on-disk format."""
    if isinstance(file, (str, os.PathLike)):
        file = os.fspath(file)
        if not file.endswith(".npz"):
            file += ".npz"
    threads = threads or min(32, len(os.sched_getaffinity(0)))
    members = []
    tasks = []
    for mi, (name, a) in enumerate(arrays.items()):
        blocks, total = _member_blocks(*_npy_parts(a))
        members.append((name + ".npy", total))
        for b, blk in enumerate(blocks):
            tasks.append((mi, (blk, b == len(blocks) - 1, level)))
    with ThreadPoolExecutor(threads) as pool:
        results = list(pool.map(_deflate_block, [t[1] for t in tasks]))
    per_member = [[] for _ in members]
    for (mi, _), res in zip(tasks, results):
        per_member[mi].append(res)
    with open(file, "wb") as fh:
        central = []
        for (name, usize), blocks in zip(members, per_member):
            crc = 0
            for _, c, ln in blocks:
                crc = _crc32_combine(crc, c, ln)
            csize = sum(len(b[0]) for b in blocks)
            central.append(_write_member(fh, name, blocks, crc, csize, usize))
        _write_central_directory(fh, central)


def _be_nice():
    """Lowers the priority of the calling THREAD (Linux: per-thread nice through its native id)."""
    try:
        import threading
        os.setpriority(os.PRIO_PROCESS, threading.get_native_id(), 10)
    except (AttributeError, OSError, PermissionError):
        pass


def default_level() -> int:
    """Deflate level of the reference files written by `newref`.  np.savez_compressed (the reference,
    newref_control.py:237) uses zlib's default, 6; on the arrays of a reference file -- float64 distances and null
    ratios, int32 indexes -- level 1 gives the same compression ratio to within 1 % (0.90 / 0.96 / 0.71) at 1.3x / 1.2x /
    7x the speed, and the stream is read by np.load exactly like any other.  WCX_NPZ_LEVEL=6 restores NumPy's level."""
    try:
        return max(0, min(9, int(os.environ.get("WCX_NPZ_LEVEL", "1"))))
    except ValueError:
        return 1


class AsyncNpzWriter:
    """savez_compressed in two halves: add(name, array) starts deflating the array on the pool right away (the caller
    goes on computing -- zlib and the CUDA library both release the GIL), close() waits and writes the archive.
    The arrays must not be modified between add() and close().  Same on-disk format as savez_compressed."""

    def __init__(self, file, threads: int | None = None, level: int | None = None):
        if level is None:
            level = default_level()
        if isinstance(file, (str, os.PathLike)):
            file = os.fspath(file)
            if not file.endswith(".npz"):
                file += ".npz"
        self.file = file
        self.level = level
        # two cores stay free for the caller, whose host work (NumPy, LAPACK) runs while the pool deflates
        # background work: the deflate threads run at a lower scheduling priority than the caller, whose host work
        # (NumPy, LAPACK, the next pass's preparation) shares the cores with them
        self.pool = ThreadPoolExecutor(threads or max(1, min(32, len(os.sched_getaffinity(0)) - 2)), initializer=_be_nice)
        # members go to the file in add() order as soon as their blocks are deflated: a writer thread follows the
        # queue, so that close() only has the tail of the last pass and the central directory left
        import queue
        import threading
        self._q = queue.Queue()
        self._central = []
        self._error = None
        self._fh = open(self.file, "wb")
        self._thread = threading.Thread(target=self._drain, name="wcx-npz-writer", daemon=True)
        self._thread.start()

    def _drain(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            if self._error is not None:
                continue  # keep consuming so that close() never blocks
            try:
                name, usize, futs, _keep = item
                blocks = [f.result() for f in futs]
                crc = 0
                for _, c, ln in blocks:
                    crc = _crc32_combine(crc, c, ln)
                csize = sum(len(b[0]) for b in blocks)
                self._central.append(_write_member(self._fh, name, blocks, crc, csize, usize))
            except BaseException as e:  # re-raised by close()
                self._error = e

    def add(self, name, array):
        header, data = _npy_parts(array)
        blocks, total = _member_blocks(header, data)
        futs = [self.pool.submit(_deflate_block, (blk, b == len(blocks) - 1, self.level)) for b, blk in enumerate(blocks)]
        self._q.put((name + ".npy", total, futs, (array, data)))

    def close(self):
        try:
            self._q.put(None)
            self._thread.join()
            if self._error is not None:
                raise self._error
            _write_central_directory(self._fh, self._central)
        finally:
            self._fh.close()
            self.pool.shutdown(wait=True)


def _write_member(fh, name, blocks, crc, csize, usize):
    offset = fh.tell()
    fname = name.encode("utf-8")
    zip64 = csize >= 0xFFFFFFFF or usize >= 0xFFFFFFFF
    extra = struct.pack("<HHQQ", 1, 16, usize, csize) if zip64 else b""
    extra += _block_index_extra(blocks)
    ver = 45 if zip64 else 20
    fh.write(struct.pack("<4sHHHHHLLLHH", b"PK\x03\x04", ver, 0, 8, 0, 0x21, crc,
                         0xFFFFFFFF if zip64 else csize, 0xFFFFFFFF if zip64 else usize, len(fname), len(extra)))
    fh.write(fname)
    fh.write(extra)
    for data, _, _ in blocks:
        fh.write(data)
    return (fname, crc, csize, usize, offset, _block_index_extra(blocks))


BLOCK_INDEX_ID = 0x6377  # private zip extra field ("wc"): compressed / uncompressed size of every independently
                         # deflated block of the member -- other readers skip unknown extra fields


def _block_index_extra(blocks):
    if len(blocks) < 2 or len(blocks) > 4000:
        return b""
    body = b"".join(struct.pack("<LL", len(data), ulen) for data, _, ulen in blocks)
    return struct.pack("<HH", BLOCK_INDEX_ID, len(body)) + body


def _write_central_directory(fh, central):
    start = fh.tell()
    for fname, crc, csize, usize, offset, bidx in central:
        zip64 = csize >= 0xFFFFFFFF or usize >= 0xFFFFFFFF or offset >= 0xFFFFFFFF
        extra = b""
        if zip64:
            extra = struct.pack("<HHQQQ", 1, 24, usize, csize, offset)
        extra += bidx
        ver = 45 if zip64 else 20
        fh.write(struct.pack("<4sHHHHHHLLLHHHHHLL", b"PK\x01\x02", ver, ver, 0, 8, 0, 0x21, crc,
                             0xFFFFFFFF if zip64 else csize, 0xFFFFFFFF if zip64 else usize, len(fname), len(extra), 0, 0, 0,
                             0o600 << 16, 0xFFFFFFFF if zip64 else offset))
        fh.write(fname)
        fh.write(extra)
    size = fh.tell() - start
    n = len(central)
    if start >= 0xFFFFFFFF or n >= 0xFFFF:
        z64 = fh.tell()
        fh.write(struct.pack("<4sQHHLLQQQQ", b"PK\x06\x06", 44, 45, 45, 0, 0, n, n, size, start))
        fh.write(struct.pack("<4sLQL", b"PK\x06\x07", 0, z64, 1))
        fh.write(struct.pack("<4sHHHHLLH", b"PK\x05\x06", 0, 0, min(n, 0xFFFF), min(n, 0xFFFF), min(size, 0xFFFFFFFF), 0xFFFFFFFF, 0))
    else:
        fh.write(struct.pack("<4sHHHHLLH", b"PK\x05\x06", 0, 0, n, n, size, start, 0))


# ---- crc32(A || B) from crc32(A), crc32(B), len(B): zlib's crc32_combine (GF(2) matrix squaring) ----
def _gf2_times(mat, vec):
    s = 0
    i = 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[i]) for i in range(32)]


def _gf2_mul(a, b):
    """matrix product a . b (columns of b mapped through a)"""
    return [_gf2_times(a, b[i]) for i in range(32)]


_CRC_SHIFT = {}  # len2 -> operator that advances a crc over len2 zero bytes


def _crc_shift_matrix(len2):
    """All blocks of a member but the last have one length, so the operator is built once per distinct length (the
    squaring chain is ~5 ms of pure Python; per block it added up to most of the final write of a 1 GB reference)."""
    m = _CRC_SHIFT.get(len2)
    if m is None:
        p = [0xEDB88320] + [1 << i for i in range(31)]  # one zero bit
        for _ in range(3):
            p = _gf2_square(p)                          # one zero byte
        m = [1 << i for i in range(32)]                 # identity
        n = len2
        while n:
            if n & 1:
                m = _gf2_mul(p, m)
            n >>= 1
            if n:
                p = _gf2_square(p)
        if len(_CRC_SHIFT) < 4096:
            _CRC_SHIFT[len2] = m
    return m


def _crc32_combine(crc1, crc2, len2):
    if len2 <= 0:
        return crc1
    return _gf2_times(_crc_shift_matrix(len2), crc1) ^ crc2


def _read_members(path, wanted):
    """{name: array} of a SMALL .npz (a sample file: ~1 MB, three members) without the zipfile module: one read of the
    file, the central directory parsed with struct, zlib.decompress per member (which releases the GIL; zipfile's
    Python-level read loop does not, and 500 sample files were 1 s of it).  Returns None for anything unusual
    (zip64, encryption, unknown method): the caller falls back to np.load."""
    with open(path, "rb") as fh:
        buf = fh.read()
    eocd = buf.rfind(b"PK\x05\x06", max(0, len(buf) - 65557))
    if eocd < 0 or eocd + 22 > len(buf):
        return None
    n, size, start = struct.unpack("<HLL", buf[eocd + 10:eocd + 20])
    if n == 0xFFFF or start == 0xFFFFFFFF or start + size > len(buf):
        return None
    out, pos = {}, start
    for _ in range(n):
        if buf[pos:pos + 4] != b"PK\x01\x02":
            return None
        flags, method = struct.unpack("<HH", buf[pos + 8:pos + 12])
        crc, csize, usize, nlen, elen, clen = struct.unpack("<LLLHHH", buf[pos + 16:pos + 34])
        off = struct.unpack("<L", buf[pos + 42:pos + 46])[0]
        name = buf[pos + 46:pos + 46 + nlen].decode("utf-8", "replace")
        pos += 46 + nlen + elen + clen
        key = name[:-4] if name.endswith(".npy") else name
        if key not in wanted:
            continue
        if (flags & 1) or csize == 0xFFFFFFFF or usize == 0xFFFFFFFF or off == 0xFFFFFFFF or buf[off:off + 4] != b"PK\x03\x04":
            return None
        lnlen, lelen = struct.unpack("<HH", buf[off + 26:off + 30])
        raw = memoryview(buf)[off + 30 + lnlen + lelen: off + 30 + lnlen + lelen + csize]
        if method == zipfile.ZIP_DEFLATED:
            data = zlib.decompress(raw, -15, usize)
        elif method == zipfile.ZIP_STORED:
            data = bytes(raw)
        else:
            return None
        if len(data) != usize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
            raise zipfile.BadZipFile("bad CRC-32 or size for member {} of {}".format(name, path))
        out[key] = np.lib.format.read_array(io.BytesIO(data), allow_pickle=True,
                                            pickle_kwargs={"encoding": "latin1", "fix_imports": True})
    return out if len(out) == len(wanted) else None


def load_samples(paths, threads: int | None = None):
    """[(sample dict, binsize)] of the sample .npz files `paths` (reference main.py:62-64), read concurrently."""
    threads = threads or min(32, len(os.sched_getaffinity(0)))

    def one(pth):
        fast = _read_members(pth, ("sample", "binsize"))
        if fast is not None:
            return fast["sample"].item(), int(fast["binsize"])
        with np.load(pth, encoding="latin1", allow_pickle=True) as npz:
            return npz["sample"].item(), int(npz["binsize"])

    with ThreadPoolExecutor(threads) as pool:
        return list(pool.map(one, paths))


def load_npz(path, threads: int | None = None):
    """dict(np.load(path, allow_pickle=True)) with the members inflated concurrently (the reference re-inflates a
    member on every `ref_file[key]` access, predict_tools.py:75,117-118,153; a 15 kb reference holds 2.5 GB)."""
    threads = threads or min(32, len(os.sched_getaffinity(0)))
    with zipfile.ZipFile(path) as zf:
        infos = zf.infolist()
    with open(path, "rb") as fh:
        fd = fh.fileno()

        def one(info):
            # local header: 30 bytes + name + extra, then the raw deflate stream
            hdr = os.pread(fd, 30, info.header_offset)
            nlen, elen = struct.unpack("<HH", hdr[26:30])
            raw = os.pread(fd, info.compress_size, info.header_offset + 30 + nlen + elen)
            bidx = _find_block_index(info.extra)
            if info.compress_type == zipfile.ZIP_DEFLATED and bidx:
                # written by savez_compressed above: the blocks are independent raw deflate segments
                data = np.empty(info.file_size, dtype=np.uint8)  # no zero fill; np.copyto releases the GIL for the block copies
                view, rawv = data, memoryview(raw)
                jobs, co, uo = [], 0, 0
                for clen, ulen in bidx:
                    jobs.append((co, clen, uo, ulen))
                    co += clen
                    uo += ulen
                if co != info.compress_size or uo != info.file_size:
                    raise ValueError("corrupt block index in " + info.filename)

                def blk(j):
                    c0, cl, u0, ul = j
                    out = zlib.decompressobj(-15).decompress(rawv[c0:c0 + cl])
                    if len(out) != ul:
                        raise ValueError("corrupt block in " + info.filename)
                    np.copyto(view[u0:u0 + ul], np.frombuffer(out, dtype=np.uint8))

                list(blk_pool.map(blk, jobs))
            elif info.compress_type == zipfile.ZIP_DEFLATED:
                data = zlib.decompress(raw, -15, info.file_size)
            elif info.compress_type == zipfile.ZIP_STORED:
                data = raw
            else:
                raise ValueError("unsupported zip compression in " + info.filename)
            bio = io.BytesIO(memoryview(data)[:4096].tobytes() if isinstance(data, np.ndarray) else data[:4096])  # header only
            version = np.lib.format.read_magic(bio)
            if version == (1, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_1_0(bio)
            else:
                shape, fortran, dtype = np.lib.format.read_array_header_2_0(bio)
            name = info.filename[:-4] if info.filename.endswith(".npy") else info.filename
            if dtype.hasobject:
                return name, np.lib.format.read_array(io.BytesIO(bytes(data)), allow_pickle=True)
            arr = np.frombuffer(data, dtype=dtype, offset=bio.tell(), count=int(np.prod(shape, dtype=np.int64)))
            arr = arr.reshape(shape, order="F" if fortran else "C")
            return name, arr.copy() if arr.size < 4096 else _writable(arr)

        with ThreadPoolExecutor(threads) as pool, ThreadPoolExecutor(threads) as blk_pool:
            return dict(pool.map(one, infos))


def _find_block_index(extra: bytes):
    i = 0
    while i + 4 <= len(extra):
        hid, ln = struct.unpack("<HH", extra[i:i + 4])
        if hid == BLOCK_INDEX_ID:
            body = extra[i + 4:i + 4 + ln]
            return [struct.unpack("<LL", body[q:q + 8]) for q in range(0, len(body), 8)]
        i += 4 + ln
    return None


def _writable(arr):
    """np.load returns writable arrays; frombuffer views of bytes are read-only.  One copy of a large member would cost
    as much as its inflation, so the view is kept and only flagged -- callers that write get a copy."""
    try:
        arr.flags.writeable = False
    except ValueError:
        pass
    return arr
