"""P1 simplex element_type tables at quadrature degree 3, in the raw column-major layout the
C ABI takes (what Fluidity passes from shape%n, shape%dn, shape%quadrature%weight).

Values: the 4-point (triangle) and 5-point (tetrahedron) degree-3 rules of Stroud (1971) as
tabulated in femtools/Quadrature.F90:690-708,951-970, with the point order produced by
expand_quadrature_template (:572-605); P1 shape functions n(i,g) = l(g,i)
(femtools/Elements.F90:511-513) and dn(i,g,k) = delta_ik, dn(loc,g,k) = -1 (:615-717).
tests/test_abi.py::test_product_tables_equal_oracle_tables checks them bit-for-bit against the oracle's restatement.
"""
import numpy as np


def quadrature_degree3(dim):
    """(l (ngi, loc), weight (ngi))"""
    if dim == 3:
        a = 0.166666666666666666666666666666666
        b = 1.0 - 3.0 * a
        l = np.array([[0.25] * 4, [a, a, a, b], [a, a, b, a], [a, b, a, a], [b, a, a, a]])
        w = np.array([-0.133333333333333333333333333333333] + [0.075] * 4)
    elif dim == 2:
        a = 0.2
        b = 1.0 - 2.0 * a
        t = 0.333333333333333333333333333333333
        l = np.array([[t] * 3, [a, a, b], [a, b, a], [b, a, a]])
        w = np.array([-0.28125] + [0.260416666666666666666666666666666] * 3)
    else:
        raise ValueError("dim must be 2 or 3")
    return l, w


def p1_tables(dim):
    """n (loc*ngi), dn (loc*ngi*dim), weight (ngi): n[i + loc*g], dn[i + loc*(g + ngi*k)]."""
    loc = dim + 1
    l, w = quadrature_degree3(dim)
    ngi = len(w)
    n = np.zeros(loc * ngi)
    dn = np.zeros(loc * ngi * dim)
    for g in range(ngi):
        for i in range(loc):
            n[i + loc * g] = l[g, i]
            for k in range(dim):
                dn[i + loc * (g + ngi * k)] = (1.0 if i == k else 0.0) if i < dim else -1.0
    return n, dn, w


def face_quadrature_degree3(dim):
    """(l (sngi, sloc), weight (sngi)) of the faces of a dim-dimensional P1 simplex mesh: the rule of
    the mesh's degree on the (dim-1)-simplex (femtools/Fields_Allocates.F90:1302-1320). dim 3: the
    triangle rule above; dim 2: the 3-point degree-3 interval rule, constants as truncated in
    femtools/Quadrature.F90:1181-1199, point order of interval_permutations (:1893-1905)."""
    if dim == 3:
        return quadrature_degree3(2)
    if dim != 2:
        raise ValueError("dim must be 2 or 3")
    a = 0.887298334620742
    b = 1.0 - a
    l = np.array([[a, b], [b, a], [0.5, 0.5]])
    w = np.array([0.277777777777777, 0.277777777777777, 0.444444444444444])
    return l, w


def p1_face_tables(dim):
    """n_f (sloc*sngi), dn_f (sloc*sngi*(dim-1)), weight_f (sngi) of mesh%faces%shape, raw column-major:
    n_f[i + sloc*g], dn_f[i + sloc*(g + sngi*k)] (P1: dn(i,g,k) = delta_ik, last node -1)."""
    sloc, fd = dim, dim - 1
    l, w = face_quadrature_degree3(dim)
    sngi = len(w)
    n = np.zeros(sloc * sngi)
    dn = np.zeros(sloc * sngi * fd)
    for g in range(sngi):
        for i in range(sloc):
            n[i + sloc * g] = l[g, i]
            for k in range(fd):
                dn[i + sloc * (g + sngi * k)] = (1.0 if i == k else 0.0) if i < fd else -1.0
    return n, dn, w
