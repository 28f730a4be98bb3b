"""Butcher tableaux: the reference's own known-answer tests (src/butcher_tableau.rs:433-502),
every literal of butcher_tableau.rs when /root/reference is mounted, and agreement between the
oracle's table, the device library's table and the generator."""
import os
import re
import subprocess
import sys

import pytest

import oracle
from ode_solvers_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/butcher_tableau.rs"


def t(w, i, j=0):
    return oracle.lib().ode_oracle_tableau(w.encode(), i, j)


def test_reference_known_answers():
    # dopri5_a/c/d/e, butcher_tableau.rs:433-455
    assert t("a5", 3, 2) == 9.0 / 40.0
    assert t("a5", 5, 3) == 64448.0 / 6561.0
    assert t("c5", 1) == 0.0 and t("c5", 3) == 3.0 / 10.0
    assert t("d5", 3) == 87487479700.0 / 32700410799.0
    assert t("d5", 7) == 69997945.0 / 29380423.0
    assert t("e5", 1) == 71.0 / 57600.0 and t("e5", 7) == -1.0 / 40.0
    # dopri853_*, butcher_tableau.rs:457-502
    assert t("a8", 2, 1) == 5.26001519587677318785587544488E-2
    assert t("a8", 7, 4) == 1.70252211019544039314978060272E-1
    assert t("a8", 12, 9) == -8.87285693353062954433549289258E0
    assert t("a8", 15, 3) == 0.0
    assert t("a8", 4, 3) == 8.87627564304205475450678981324E-2
    assert t("b8", 6) == 4.45031289275240888144113950566E0
    assert t("b8", 10) == -1.52160949662516078556178806805E-1
    assert t("c8", 3) == 0.789002279381515978178381316732E-01
    assert t("c8", 8) == 0.307692307692307692307692307692E+00
    assert t("d8", 6, 4) == 0.0
    assert t("d8", 5, 9) == -0.22113666853125306036270938578E+02
    assert t("e8", 1) == 0.1312004499419488073250102996E-01
    assert t("e8", 12) == -0.2235530786388629525884427845E-01
    assert t("bhh8", 1) == 0.244094488188976377952755905512E+00
    # the reference's quirk (Q1): c12 = c13 = 0.0
    assert t("c8", 12) == 0.0 and t("c8", 13) == 0.0


def test_device_table_equals_oracle_table():
    lib = _abi.load()
    d = lambda w, i, j=0: lib.ode_b200_tableau(w.encode(), i, j)  # noqa: E731  (0-based)
    for i in range(2, 8):
        for j in range(1, i):
            assert d("a5", i - 1, j - 1) == t("a5", i, j)
    for i in range(1, 8):
        assert d("c5", i - 1) == t("c5", i) and d("d5", i - 1) == t("d5", i) and d("e5", i - 1) == t("e5", i)
    for i in range(2, 17):
        for j in range(1, i):
            assert d("a8", i - 1, j - 1) == t("a8", i, j)
    for i in range(1, 13):
        assert d("b8", i - 1) == t("b8", i)
    for i in range(1, 4):
        assert d("bhh8", i - 1) == t("bhh8", i)
    for i in range(1, 17):
        assert d("c8", i - 1) == t("c8", i) and d("e8", i - 1) == t("e8", i)
    for i in range(4, 8):
        for j in range(1, 17):
            assert d("d8", i - 4, j - 1) == t("d8", i, j)


def test_generated_headers_are_current(tmp_path):
    """tools/gen_tableau.py reproduces the committed headers byte for byte."""
    for rel in ("oracle/ode_tableau_c.h", "ode_solvers_b200/csrc/ode_tableau.cuh"):
        before = open(os.path.join(ROOT, rel)).read()
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tableau.py")], check=True,
                       stdout=subprocess.DEVNULL)
        assert open(os.path.join(ROOT, rel)).read() == before


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not mounted (GPU box)")
def test_every_literal_of_the_reference():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_tableau as g
    src = open(REF).read().split("#[cfg(test)]")[0]
    d5, d8 = src.split("pub(crate) mod dopri853")

    def lits(block):
        out = []
        for m in re.finditer(r"(-?\d[\d\.]*(?:E[+-]?\d+)?)(?:\s*/\s*(\d[\d\.]*))?\s*[,\]\n]", block):
            a = float(m.group(1))
            out.append(a / float(m.group(2)) if m.group(2) else a)
        return out

    def const(block, name):
        return lits(re.search(r"const %s:.*?= (\(|\[)(.*?);\n" % name, block, re.S).group(2))

    assert const(d5, "A") == [v for i in range(2, 8) for v in g.DP5_A[i]]
    assert const(d5, "C") == g.DP5_C and const(d5, "D") == g.DP5_D and const(d5, "E") == g.DP5_E
    a8 = lits(re.search(r"\) = \((.*?)\n    \);", d8, re.S).group(1))
    assert a8 == [float(v) for i in range(2, 17) for v in g.DP8_A_TXT[i]] and len(a8) == 120
    assert const(d8, "B") == [float(v) for v in g.DP8_B_TXT]
    assert const(d8, "BHH") == [float(v) for v in g.DP8_BHH_TXT]
    assert const(d8, "C") == [float(v) for v in g.DP8_C_TXT]
    assert const(d8, "D") == [float(v) for r in g.DP8_D_TXT for v in r]
    assert const(d8, "E") == [float(v) for v in g.DP8_E_TXT]
