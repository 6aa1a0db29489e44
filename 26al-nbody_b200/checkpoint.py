"""State checkpoints and resume (/root/reference/al26_nbody.py:83-123 `Metadata`, :281-293 `State`,
:295-318 `most_recent_checkpoint`, :320-345 `compress` / `decompress`, :347-401 `save_checkpoint`, :403-439
`load_checkpoint`; resume in `main`, :1641-1656 and :1734-1737).

Same files as the reference writes, so its plotting scripts find them:
    <base>-state-NNNNN.pkl.zst     pickle of State(cluster, converter, metadata), zstd level 8
    <base>-yields.ubj.zst          the Yields book (yields_io.Yields.marinate)
`zstandard` / `ubjson` are not installable in every environment (they are absent from this image); without them the
state goes to `<base>-state-NNNNN.pkl.gz` (gzip) and the yields to `<base>-yields.ubj.zst.npz` -- `load_checkpoint` reads
whichever exists.  Host-side I/O, not part of the hot path: the device-resident inventories are pulled into the cluster
columns on save steps (driver.pull_inventories) and pushed back on resume (driver.push_inventories).

What a checkpoint does NOT hold -- in the reference either: the gravity worker's forces and
individual timesteps.  After a reload the worker is re-created from (mass, position, velocity), its clock is set with
`gravity.model_time = metadata.time` (:1736) and the first evolve re-initialises forces and timesteps.
"""
import glob
import gzip
import os
import pickle
import re
from datetime import datetime

from . import units as U


class Metadata:
    """al26_nbody.py:83-123 (the argparse namespace is optional here)."""

    def __init__(self, args=None, t_f=None, model="plummer", nstars=0, cluster_radius=None, filename=""):
        self.sim_start = datetime.now()
        self.sim_start_str = self.sim_start.strftime("%d/%m/%Y %H:%M:%S")
        self.update_access_time()
        self.args = args
        self.model = getattr(args, "model", model)
        self.nstars = getattr(args, "n", nstars)
        self.cluster_radius = getattr(args, "rc", cluster_radius)
        fn = getattr(args, "filename", filename)
        self.filename = fn if fn != "" else self.generate_filename()
        self.time = 0.0 | U.Myr
        self.t_f = t_f
        self.completion = 0.0
        self.most_recent_checkpoint = 0

    def generate_filename(self):
        return "sim-" + self.sim_start.strftime("%Y-%m-%d-%H-%M-%S")

    def update(self, current_time, increment_checkpoint=True):
        if increment_checkpoint:
            self.most_recent_checkpoint += 1
        self.update_completion(current_time)
        self.update_access_time()

    def update_completion(self, current_time):
        self.time = current_time
        if self.t_f is not None:
            self.completion = float(U.value_in(self.time, U.Myr)) / float(U.value_in(self.t_f, U.Myr))

    def update_access_time(self):
        self.sim_last = datetime.now()
        self.sim_last_str = self.sim_last.strftime("%d/%m/%Y %H:%M:%S")


class State:
    """al26_nbody.py:281-293: what one state file holds."""

    def __init__(self, cluster, converter, metadata):
        self.cluster = cluster
        self.converter = converter
        self.metadata = metadata


def _zstd():
    try:
        import zstandard
        return zstandard
    except ImportError:
        return None


def compress(data, level=8, threads=-1):
    """al26_nbody.py:320-333 (zstd); gzip where zstandard is absent.  Returns (bytes, extension)."""
    z = _zstd()
    if z is not None:
        return z.ZstdCompressor(threads=threads, level=level).compress(data), ".zst"
    return gzip.compress(data, compresslevel=6), ".gz"


def decompress(data, ext):
    if ext == ".zst":
        z = _zstd()
        if z is None:
            raise IOError("this checkpoint is zstd-compressed and the `zstandard` module is not installed")
        return z.ZstdDecompressor().decompress(data)
    return gzip.decompress(data)


def _state_path(filename, nfile):
    """the state file of checkpoint nfile, whichever compression it was written with (None if there is none)"""
    stem = filename + "-state-" + str(nfile).zfill(5) + ".pkl"
    for ext in (".zst", ".gz"):
        if os.path.isfile(stem + ext):
            return stem + ext, ext
    return None, None


def most_recent_checkpoint(filename):
    """al26_nbody.py:295-318: the highest NNNNN among <base>-state-NNNNN.*"""
    rx = re.compile(r"-state-(\d+)\.pkl")
    highest = -1
    for f in glob.glob(glob.escape(filename) + "-state-*"):
        m = rx.search(f)
        if m:
            highest = max(highest, int(m.group(1)))
    if highest < 0 or _state_path(filename, highest)[0] is None:
        raise IOError("Missing file! Somethings up!")
    return highest


def save_checkpoint(filename, nfile, cluster, converter, yields, metadata):
    """al26_nbody.py:347-401.  Returns (state path, yields path)."""
    blob, ext = compress(pickle.dumps(State(cluster, converter, metadata)))
    state_filename = filename + "-state-" + str(nfile).zfill(5) + ".pkl" + ext
    with open(state_filename, "wb") as f:
        f.write(blob)
    yields_path = yields.marinate(filename + "-yields.ubj.zst") if yields is not None else None
    return state_filename, yields_path


def load_checkpoint(filename, nfile):
    """al26_nbody.py:403-439.  Returns (cluster, converter, yields, metadata); yields is None when no book was written."""
    from .yields_io import Yields
    path, ext = _state_path(filename, nfile)
    if path is None:
        raise IOError("no state file for checkpoint {} of {}".format(nfile, filename))
    with open(path, "rb") as f:
        state = pickle.loads(decompress(f.read(), ext))
    yields = None
    for yp in (filename + "-yields.ubj.zst", filename + "-yields.ubj.zst.npz"):
        if os.path.isfile(yp):
            yields = Yields(filename)
            yields.plate(yp)
            yields.first_write = False  # the CSV already has its header
            break
    return state.cluster, state.converter, yields, state.metadata
