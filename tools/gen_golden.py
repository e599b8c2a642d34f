#!/usr/bin/env python
"""Generate golden vectors under tests/golden/ by importing the Python parts of
the reference that import cleanly in this container:
``w2dyn.dmft.interaction`` (U-matrices, auto quantum numbers) and
``w2dyn.auxiliaries.transform`` (hybr_from_sites, leg2mat).

Run here only (needs /root/reference); the outputs are committed.
    python tools/gen_golden.py
"""
import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
import w2dyn.dmft.interaction as interaction     # noqa: E402
import w2dyn.auxiliaries.transform as tf         # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sparse(um):
    um = np.asarray(um)
    if np.iscomplexobj(um):
        assert np.abs(um.imag).max() < 1e-12
        um = um.real
    idx = np.argwhere(um != 0.0).astype(np.int16)
    val = um[um != 0.0]
    return idx, val


def main():
    # --- U matrices of the five configs (interaction.py:171-425)
    items = {
        "density_1": interaction.Density(1, 2.0, 0.0, 0.0),
        "density_2": interaction.Density(2, 4.0, 2.5, 0.75),
        "kanamori_3": interaction.Kanamori(3, 4.0, 2.8, 0.6),
        "coulomb_5": interaction.Coulomb(5, 4.0, 8.615, 5.385, 0.0),
        "coulomb_3": interaction.Coulomb(3, 3.0, 5.0, 0.0, 0.0),
    }
    store = {}
    for name, it in items.items():
        idx, val = sparse(it.u_matrix)
        store[name + "_idx"] = idx
        store[name + "_val"] = val
        store[name + "_qns"] = np.array(" ".join(it.auto_quantum_numbers))
        store[name + "_udens"] = np.asarray(it.get_udens()).real
    np.savez_compressed(os.path.join(OUT, "umatrix.npz"), **store)

    # --- hybridisation from bath sites (transform.py:114-146), small grid
    beta = 10.0
    nb = 8
    p = np.arange(1, nb + 1)
    eps = 2.0 * np.cos(p * np.pi / (nb + 1))
    v = np.sqrt((2.0 / (nb + 1)) * np.sin(p * np.pi / (nb + 1)) ** 2)
    tau = np.linspace(0, beta, 51)
    iw = tf.matfreq(beta, "fermi", 20)
    hiw, htau = tf.hybr_from_sites(eps, v[:, None], iw, tau, beta)
    # leg2mat for NLegOrder=30, 20 frequencies (transform.py:221-255)
    l2m = tf.leg2mat(30, 20)
    np.savez_compressed(os.path.join(OUT, "hybr_leg.npz"),
                        beta=beta, eps=eps, v=v, tau=tau, iw=iw,
                        hybr_tau=htau[0, 0].real, hybr_iw=hiw[0, 0], leg2mat=l2m)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
