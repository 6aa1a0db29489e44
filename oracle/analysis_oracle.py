"""oracle/analysis_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the O(N^2) 10-nearest-neighbour local density of the reference's post-processing
(`local_densities_numba`, /root/reference/plotting/al26_plot.py:324-359).  Pinned bit-for-bit
against the reference's own numba function lifted in the build container (oracle/lift_reference.py ->
tests/golden/enrich_golden.npz, keys ld_*).
"""
import numpy as np

FTP = 4.18879020479  # the script's literal four-thirds pi (:326)


def local_densities(x, y, z, masses):
    n = len(x)
    rho = np.zeros(n)
    for i in range(n):
        dx, dy, dz = x[i] - x, y[i] - y, z[i] - z
        d = np.sqrt(dx * dx + dy * dy + dz * dz)          # (xi-xj)**2 + (yi-yj)**2 + (zi-zj)**2, sqrt
        order = np.argsort(d, kind="stable")
        nr = order[1:11]                                   # 10 nearest (index 0 is the star itself)
        d10 = d[nr[-1]]
        mass = 0.0
        for j in nr:                                       # summed in ascending-distance order
            mass += masses[j]
        rho[i] = mass / (FTP * d10 * d10 * d10)
    return rho
