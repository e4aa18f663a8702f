"""A synthetic LARGE constraint DAG (SURVEY 7 "Dynamic DAGs": the reference's BLAKE3-compression circuits have thousands of
nodes; this is the stress case for the device interpreter's 1024 / 4096-slot instantiations). Built with the independent Python
compiler of tests/_pyverifier.py and handed to both sides as msgpu_graph_desc descriptors."""
from tests import _pyverifier as pv


def big_dag_circuit(width=48, k=1200, n_lookups=4):
    """k degree-2 products P_i = col_a * col_b + (i + 1); constraint set A pairs P_j with P_{j + k/2}, set B pairs P_j with
    P_{k-1-j}: every product is used twice, far apart in the node order, so about k values are live at once (the lowering
    cannot recycle their slots) -- k = 1200 needs the 4096-slot interpreter, k = 300 the 1024-slot one. Degree 3, so q = 2."""
    E = pv.Expr
    m = E.main
    P = [m(i % width) * m((7 * i + 3) % width) + E.const(i + 1) for i in range(k)]
    cons = [P[j] * m((5 * j) % width) - P[j + k // 2] * m((11 * j + 1) % width) for j in range(k // 2)]
    cons += [P[j] * m((3 * j + 2) % width) - P[k - 1 - j] for j in range(k // 2)]
    lookups = []
    for t in range(n_lookups):
        args = [E.const(t), m(t) + m(t + 1) * E.const(256), m(t + 2) * m(t + 3)]
        lookups.append(pv.push(m(width - 1 - t), args) if t % 2 == 0 else pv.pull(m(width - 1 - t), args))
    return pv.Circuit(width, lookups, cons)
