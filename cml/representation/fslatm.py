"""f2py module name of coreml/cml/representation/fslatm.f90 (same positional signatures)."""
import numpy as np

from aqml_b200 import _lib as L


def _call(fn, coordinates, nuclear_charges, ia, izeff, zs_type, tail):
    coords = L.f64(coordinates)
    zs = L.i32(np.asarray(nuclear_charges))
    nx = tail[1]
    ys = np.zeros(nx)
    L.call(fn, L.ptr(coords), L.ptr(zs), len(zs), int(ia), int(bool(izeff)), *[int(z) for z in zs_type], *tail, L.ptr(ys))
    return ys


def calc_sbop(coordinates, nuclear_charges, cg, izeff, z1, z2, rcut, nx, dgrid, sigma, coeff, rpower):
    return _call("aqml_calc_sbop", coordinates, nuclear_charges, -1, izeff, (z1, z2),
                 (float(rcut), int(nx), float(dgrid), float(sigma), float(coeff), float(rpower)))


def calc_sbop_local(coordinates, nuclear_charges, ia_python, cg, izeff, z1, z2, rcut, nx, dgrid, sigma, coeff, rpower):
    return _call("aqml_calc_sbop", coordinates, nuclear_charges, ia_python, izeff, (z1, z2),
                 (float(rcut), int(nx), float(dgrid), float(sigma), float(coeff), float(rpower)))


def calc_sbot(coordinates, nuclear_charges, cg, izeff, z1, z2, z3, rcut, nx, dgrid, sigma, coeff):
    return _call("aqml_calc_sbot", coordinates, nuclear_charges, -1, izeff, (z1, z2, z3),
                 (float(rcut), int(nx), float(dgrid), float(sigma), float(coeff)))


def calc_sbot_local(coordinates, nuclear_charges, ia_python, cg, izeff, z1, z2, z3, rcut, nx, dgrid, sigma, coeff):
    return _call("aqml_calc_sbot", coordinates, nuclear_charges, ia_python, izeff, (z1, z2, z3),
                 (float(rcut), int(nx), float(dgrid), float(sigma), float(coeff)))
