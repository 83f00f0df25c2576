"""Host mirror of the reference's archive layer over the C ABI.

  ArchiveWriter(input_dir, output_path).pack(files)   reference src/archive/writer.rs:66-195
  ArchiveReader(path).get_summary() / .unpack(dir)    reference src/archive/reader.rs:46-244
  ArchiveSummary / FileEntry                           reference src/archive/reader.rs:25-38
  walk_dir                                             reference src/fsutil/directory.rs:39-73
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

from . import _lib as L
from .context import Context


@dataclass
class FileEntry:
    path: str
    original_size: int


@dataclass
class ArchiveSummary:
    unique_chunks: int
    total_original_size: int
    archive_size: int
    compression_ratio: float
    squish_creation_date: int
    squish_version: str
    files: List[FileEntry] = field(default_factory=list)


def walk_dir(root) -> List[Path]:
    """Iterative stack DFS, directories via is_dir() (follows symlinks), everything else a file."""
    root = Path(root)
    if not root.is_dir():
        raise L.SquishError(-2, f"Failed to read directory {root}")
    out, stack = [], [root]
    while stack:
        d = stack.pop()
        for e in os.scandir(d):
            (stack if Path(e.path).is_dir() else out).append(Path(e.path))
    return out


class ArchiveWriter:
    def __init__(self, input_dir, output_path, ctx: Optional[Context] = None, threads: int = 25, ctxs: Optional[List[Context]] = None):
        """`ctxs`: one Context per GPU of this box (sq_archive_pack_multi); the first one owns the dedup index."""
        self.input_dir, self.output_path, self.threads = str(input_dir), str(output_path), threads
        self.ctxs = list(ctxs) if ctxs else [ctx or Context()]
        self.ctx = self.ctxs[0]
        self.report = None

    def pack(self, files=None) -> int:
        """Packs the directory tree; returns the archive size like the reference.  `files` is accepted
        for signature parity; the native packer walks input_dir itself (same walk rule)."""
        rep = L.SqPackReport()
        if len(self.ctxs) > 1:
            hs = (C.c_void_p * len(self.ctxs))(*[c.h for c in self.ctxs])
            self.ctx.check(self.ctx.lib.sq_archive_pack_multi(hs, len(self.ctxs), self.input_dir.encode(), self.output_path.encode(), self.threads, C.byref(rep)))
        else:
            self.ctx.check(self.ctx.lib.sq_archive_pack(self.ctx.h, self.input_dir.encode(), self.output_path.encode(), self.threads, C.byref(rep)))
        self.report = rep
        return rep.archive_size


class ArchiveReader:
    def __init__(self, archive_path, ctx: Optional[Context] = None, threads: int = 25):
        self.path, self.threads, self.ctx = str(archive_path), threads, ctx
        self.lib = L.load()
        self._summary = self._list()  # ArchiveReader::new validates header + index up front

    def _list(self) -> ArchiveSummary:
        s = L.SqSummary()
        p = C.c_char_p()
        rc = self.lib.sq_archive_list(self.path.encode(), C.byref(s), C.byref(p))
        if rc != L.SQ_OK:
            raise L.SquishError(rc, (self.lib.sq_last_error(None) or b"").decode())
        text = p.value.decode() if p.value else ""
        self.lib.sq_free(p)
        files = []
        for line in text.splitlines():
            size, _, path = line.partition(" ")
            files.append(FileEntry(path, int(size)))
        return ArchiveSummary(s.unique_chunks, s.total_original_size, s.archive_size, s.compression_ratio, s.timestamp,
                              s.version.decode(), files)

    def get_summary(self) -> ArchiveSummary:
        return self._summary

    def unpack(self, output_dir, ctxs: Optional[List[Context]] = None) -> None:
        """`ctxs`: one Context per GPU of this box (sq_archive_unpack_multi: records split into contiguous ranges)."""
        ctx = ctxs[0] if ctxs else (self.ctx or Context())
        s = L.SqSummary()
        if ctxs and len(ctxs) > 1:
            hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
            ctx.check(ctx.lib.sq_archive_unpack_multi(hs, len(ctxs), self.path.encode(), str(output_dir).encode(), self.threads, C.byref(s)))
        else:
            ctx.check(ctx.lib.sq_archive_unpack(ctx.h, self.path.encode(), str(output_dir).encode(), self.threads, C.byref(s)))
        self.unpack_summary = s
