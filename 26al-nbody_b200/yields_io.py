"""Yields bookkeeping and on-disk formats, mirroring the reference's `Yields` class
(/root/reference/al26_nbody.py:125-279) so its plotting scripts keep working:

  <base>-cluster-yields.csv   header `time,local_26al,global_26al,sne_26al,local_60fe,global_60fe,sne_60fe`,
                              one `{:.6e}` row of cluster sums [Msun] per save step (:221-241)
  <base>-yields.*             the whole object as a dict with the reference's attribute names (:242-264):
                              UBJSON + zstd when `ubjson` and `zstandard` are importable (the reference's
                              format), otherwise a compressed .npz with the same keys.

Host-side I/O, not part of the hot path: values arrive from the device through `driver.pull_inventories`.
"""
import numpy as np

from . import units as U

_SERIES = ("local_26al", "global_26al", "sne_26al", "agb_26al", "agb_26al_raw",
           "local_60fe", "global_60fe", "sne_60fe", "agb_60fe", "agb_60fe_raw")
_SUMS = ("sum_local_26al", "sum_global_26al", "sum_sne_26al", "sum_agb_26al",
         "sum_local_60fe", "sum_global_60fe", "sum_sne_60fe", "sum_agb_60fe")
_FINALS = ("local_26al_final", "global_26al_final", "sne_26al_final", "agb_26al_final",
           "local_60fe_final", "global_60fe_final", "sne_60fe_final", "agb_60fe_final")


def _column(cluster, series):
    """'local_26al' -> cluster.mass_26al_local, 'agb_60fe_raw' -> cluster.mass_60fe_agb_raw,
    'sne_26al_final' -> cluster.mass_26al_sne_final; in Msun, as a list (the reference stores lists)."""
    parts = series.split("_")
    name = "mass_{}_{}".format(parts[1], "_".join([parts[0]] + parts[2:]))
    col = getattr(cluster, name, None)
    if col is None:
        return [0.0] * len(cluster)
    return list(np.asarray(U.value_in(col, U.MSun), dtype=np.float64))


class Yields:
    def __init__(self, filename):
        self.filename = filename
        self.time = []
        for k in _SERIES + _SUMS + _FINALS:
            setattr(self, k, [])
        self.first_write = True

    def update_state(self, model_time, cluster):
        """al26_nbody.py:169-220"""
        self.time.append(float(U.value_in(model_time, U.Myr)))
        for k in _SERIES:
            getattr(self, k).append(_column(cluster, k))
        for k in _SUMS:
            getattr(self, k).append(sum(getattr(self, k[4:])[-1]))  # python sum over the list, as the reference
        for k in _FINALS:
            setattr(self, k, _column(cluster, k))
        if self.first_write:
            self.write_csv_header()
            self.first_write = False
        self.write_to_csv()

    def write_csv_header(self):
        with open("{}-cluster-yields.csv".format(self.filename), "w") as f:
            f.write("time,local_26al,global_26al,sne_26al,local_60fe,global_60fe,sne_60fe\n")

    def write_to_csv(self):
        with open("{}-cluster-yields.csv".format(self.filename), "a") as f:
            f.write("{:.6e},{:.6e},{:.6e},{:.6e},{:.6e},{:.6e},{:.6e}\n".format(
                self.time[-1], self.sum_local_26al[-1], self.sum_global_26al[-1], self.sum_sne_26al[-1],
                self.sum_local_60fe[-1], self.sum_global_60fe[-1], self.sum_sne_60fe[-1]))

    def _dict(self):
        return {attr: value for attr, value in self.__dict__.items()}

    def marinate(self, filename):
        """Serialise the whole object (:242-264).  Returns the path written."""
        try:
            import ubjson
            import zstandard
            data = zstandard.ZstdCompressor(level=8, threads=-1).compress(ubjson.dumpb(self._dict()))
            with open(filename, "wb") as f:
                f.write(data)
            return filename
        except ImportError:
            path = filename if filename.endswith(".npz") else filename + ".npz"
            np.savez_compressed(path, **{k: np.asarray(v) if not isinstance(v, (str, bool)) else np.asarray(v)
                                         for k, v in self._dict().items()})
            return path

    def plate(self, filename):
        """inverse of marinate (:265-279)"""
        if filename.endswith(".npz"):
            z = np.load(filename, allow_pickle=False)
            for attr in list(self.__dict__):
                v = z[attr]
                self.__dict__[attr] = v.item() if v.ndim == 0 else v.tolist()
            return
        import ubjson
        import zstandard
        with open(filename, "rb") as f:
            preserve = ubjson.loadb(zstandard.ZstdDecompressor().decompress(f.read()))
        for attr in list(self.__dict__):
            self.__dict__[attr] = preserve[attr]
