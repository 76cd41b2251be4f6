"""``tf.gfile``: the file API TF-1.x programs use for checkpoint / log directories, over the local filesystem.  URL-style paths map
through :func:`train.saver.resolve_path` -- the reference keeps its checkpoints under ``hdfs://...``
(``/root/reference/distributed_mnist.py:127``, ``example_between_graph.py:91``); here ``DTF_HDFS_ROOT`` plays the shared filesystem --
so ``tf.gfile.Exists(checkpoint_dir)`` and ``MonitoredTrainingSession(checkpoint_dir=...)`` agree on where that is."""
from __future__ import annotations

import glob as _glob
import os
import shutil
from typing import Iterator, List, Tuple

from ..framework import errors

__all__ = ["Exists", "IsDirectory", "MakeDirs", "MkDir", "ListDirectory", "Glob", "Remove", "DeleteRecursively", "Rename", "Copy", "Stat",
           "Walk", "GFile", "FastGFile", "Open"]


def _p(path) -> str:
    from ..train.saver import resolve_path
    return resolve_path(path.decode() if isinstance(path, bytes) else str(path))


def Exists(filename) -> bool:                    # noqa: N802 - TF's names throughout
    return os.path.exists(_p(filename))


def IsDirectory(dirname) -> bool:                # noqa: N802
    return os.path.isdir(_p(dirname))


def MakeDirs(dirname) -> None:                   # noqa: N802
    os.makedirs(_p(dirname), exist_ok=True)


def MkDir(dirname) -> None:                      # noqa: N802
    try:
        os.mkdir(_p(dirname))
    except FileNotFoundError as e:
        raise errors.NotFoundError(str(e)) from None
    except FileExistsError as e:
        raise errors.OpError(str(e)) from None


def ListDirectory(dirname) -> List[str]:         # noqa: N802
    try:
        return sorted(os.listdir(_p(dirname)))
    except FileNotFoundError as e:
        raise errors.NotFoundError(str(e)) from None


def Glob(filename) -> List[str]:                 # noqa: N802
    return sorted(_glob.glob(_p(filename)))


def Remove(filename) -> None:                    # noqa: N802
    try:
        os.remove(_p(filename))
    except FileNotFoundError as e:
        raise errors.NotFoundError(str(e)) from None


def DeleteRecursively(dirname) -> None:          # noqa: N802
    try:
        shutil.rmtree(_p(dirname))
    except FileNotFoundError as e:
        raise errors.NotFoundError(str(e)) from None


def Rename(oldname, newname, overwrite: bool = False) -> None:      # noqa: N802
    dst = _p(newname)
    if os.path.exists(dst) and not overwrite:
        raise errors.OpError("file already exists: %s" % newname)
    os.replace(_p(oldname), dst)


def Copy(oldpath, newpath, overwrite: bool = False) -> None:        # noqa: N802
    dst = _p(newpath)
    if os.path.exists(dst) and not overwrite:
        raise errors.OpError("file already exists: %s" % newpath)
    shutil.copyfile(_p(oldpath), dst)


class _Stat:
    def __init__(self, st):
        self.length, self.mtime_nsec, self.is_directory = st.st_size, st.st_mtime_ns, os.path.stat.S_ISDIR(st.st_mode)


def Stat(filename) -> _Stat:                     # noqa: N802
    try:
        return _Stat(os.stat(_p(filename)))
    except FileNotFoundError as e:
        raise errors.NotFoundError(str(e)) from None


def Walk(top, in_order: bool = True) -> Iterator[Tuple[str, List[str], List[str]]]:      # noqa: N802
    for root, dirs, files in os.walk(_p(top), topdown=in_order):
        yield root, sorted(dirs), sorted(files)


class GFile:
    """File object with TF's method names (``read`` / ``write`` / ``readline(s)`` / ``size`` / ``flush`` / ``close``, iteration,
    ``with``); text mode unless the mode has a ``b``."""

    def __init__(self, name, mode: str = "r"):
        self.name, self.mode = name, mode
        path = _p(name)
        try:
            self._f = open(path, mode)
        except FileNotFoundError as e:
            raise errors.NotFoundError(str(e)) from None

    def read(self, n: int = -1):
        return self._f.read(n)

    def write(self, data) -> None:
        self._f.write(data)

    def readline(self):
        return self._f.readline()

    def readlines(self):
        return self._f.readlines()

    def seek(self, offset: int, whence: int = 0):
        return self._f.seek(offset, whence)

    def tell(self) -> int:
        return self._f.tell()

    def size(self) -> int:
        self._f.flush() if "r" not in self.mode or "+" in self.mode else None
        return os.path.getsize(_p(self.name))

    def flush(self) -> None:
        self._f.flush()

    def close(self) -> None:
        self._f.close()

    def __iter__(self):
        return iter(self._f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


FastGFile = GFile
Open = GFile
