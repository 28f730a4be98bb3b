#!/usr/bin/env python3
"""Generate the Butcher-tableau headers from ONE coefficient table.

Emits (both committed):
  oracle/ode_tableau_c.h                   plain C arrays for the CPU oracle
  ode_solvers_b200/csrc/ode_tableau.cuh    constexpr accessors for the sm_100a kernels

The coefficients are the published Dormand-Prince 5(4) rationals and the
Hairer/Wanner DOP853 30-digit constants, i.e. the same numbers the reference
keeps in src/butcher_tableau.rs:10-54 (dopri54) and :91-377 (dopri853),
INCLUDING its quirk C[11] = C[12] = 0.0 (butcher_tableau.rs:278-279)
and the all-zero placeholder row 13 (:194).

Every value is emitted as a C99/C++17 hex-float literal computed here with
IEEE double arithmetic (Python float): a rational p/q is float(p)/float(q),
exactly how rustc constant-folds `19372.0 / 6561.0`; a decimal literal is
rounded to nearest by Python's correctly-rounded parser, as rustc does. Hex
literals remove any dependence on how gcc / nvcc parse long decimals.

tests/test_tableau.py checks the generated numbers against the reference's
own known-answer tests (butcher_tableau.rs:433-502) and, when /root/reference
is present, against every literal in butcher_tableau.rs.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def q(p, d):
    return float(p) / float(d)


# ---------------------------------------------------------------- DOPRI5(4)
# 1-based row i = 2..7, column j = 1..i-1 (row 7 = 5th-order weights, FSAL)
DP5_A = {
    2: [q(1, 5)],
    3: [q(3, 40), q(9, 40)],
    4: [q(44, 45), q(-56, 15), q(32, 9)],
    5: [q(19372, 6561), q(-25360, 2187), q(64448, 6561), q(-212, 729)],
    6: [q(9017, 3168), q(-355, 33), q(46732, 5247), q(49, 176), q(-5103, 18656)],
    7: [q(35, 384), 0.0, q(500, 1113), q(125, 192), q(-2187, 6784), q(11, 84)],
}
DP5_C = [0.0, q(1, 5), q(3, 10), q(4, 5), q(8, 9), 1.0, 1.0]
DP5_D = [
    q(-12715105075, 11282082432), 0.0, q(87487479700, 32700410799),
    q(-10690763975, 1880347072), q(701980252875, 199316789632),
    q(-1453857185, 822651844), q(69997945, 29380423),
]
DP5_E = [
    q(71, 57600), 0.0, q(-71, 16695), q(71, 1920), q(-17253, 339200),
    q(22, 525), q(-1, 40),
]

# ---------------------------------------------------------------- DOP853
Z = "0.0"
DP8_A_TXT = {
    2: ["5.26001519587677318785587544488E-2"],
    3: ["1.97250569845378994544595329183E-2", "5.91751709536136983633785987549E-2"],
    4: ["2.95875854768068491816892993775E-2", Z, "8.87627564304205475450678981324E-2"],
    5: ["2.41365134159266685502369798665E-1", Z, "-8.84549479328286085344864962717E-1",
        "9.24834003261792003115737966543E-1"],
    6: ["3.7037037037037037037037037037E-2", Z, Z, "1.70828608729473871279604482173E-1",
        "1.25467687566822425016691814123E-1"],
    7: ["3.7109375E-2", Z, Z, "1.70252211019544039314978060272E-1",
        "6.02165389804559606850219397283E-2", "-1.7578125E-2"],
    8: ["3.70920001185047927108779319836E-2", Z, Z, "1.70383925712239993810214054705E-1",
        "1.07262030446373284651809199168E-1", "-1.53194377486244017527936158236E-2",
        "8.27378916381402288758473766002E-3"],
    9: ["6.24110958716075717114429577812E-1", Z, Z, "-3.36089262944694129406857109825E0",
        "-8.68219346841726006818189891453E-1", "2.75920996994467083049415600797E1",
        "2.01540675504778934086186788979E1", "-4.34898841810699588477366255144E1"],
    10: ["4.77662536438264365890433908527E-1", Z, Z, "-2.48811461997166764192642586468E0",
         "-5.90290826836842996371446475743E-1", "2.12300514481811942347288949897E1",
         "1.52792336328824235832596922938E1", "-3.32882109689848629194453265587E1",
         "-2.03312017085086261358222928593E-2"],
    11: ["-9.3714243008598732571704021658E-1", Z, Z, "5.18637242884406370830023853209E0",
         "1.09143734899672957818500254654E0", "-8.14978701074692612513997267357E0",
         "-1.85200656599969598641566180701E1", "2.27394870993505042818970056734E1",
         "2.49360555267965238987089396762E0", "-3.0467644718982195003823669022E0"],
    12: ["2.27331014751653820792359768449E0", Z, Z, "-1.05344954667372501984066689879E1",
         "-2.00087205822486249909675718444E0", "-1.79589318631187989172765950534E1",
         "2.79488845294199600508499808837E1", "-2.85899827713502369474065508674E0",
         "-8.87285693353062954433549289258E0", "1.23605671757943030647266201528E1",
         "6.43392746015763530355970484046E-1"],
    13: [Z] * 12,   # placeholder row (the 8th-order weights live in B)
    14: ["5.61675022830479523392909219681E-2", Z, Z, Z, Z, Z,
         "2.53500210216624811088794765333E-1", "-2.46239037470802489917441475441E-1",
         "-1.24191423263816360469010140626E-1", "1.5329179827876569731206322685E-1",
         "8.20105229563468988491666602057E-3", "7.56789766054569976138603589584E-3",
         "-8.298E-3"],
    15: ["3.18346481635021405060768473261E-2", Z, Z, Z, Z,
         "2.83009096723667755288322961402E-2", "5.35419883074385676223797384372E-2",
         "-5.49237485713909884646569340306E-2", Z, Z,
         "-1.08347328697249322858509316994E-4", "3.82571090835658412954920192323E-4",
         "-3.40465008687404560802977114492E-4", "1.41312443674632500278074618366E-1"],
    16: ["-4.28896301583791923408573538692E-1", Z, Z, Z, Z,
         "-4.69762141536116384314449447206E0", "7.68342119606259904184240953878E0",
         "4.06898981839711007970213554331E0", "3.56727187455281109270669543021E-1",
         Z, Z, Z, "-1.39902416515901462129418009734E-3",
         "2.9475147891527723389556272149E0", "-9.15095847217987001081870187138E0"],
}
DP8_B_TXT = [
    "5.42937341165687622380535766363E-2", Z, Z, Z, Z,
    "4.45031289275240888144113950566E0", "1.89151789931450038304281599044E0",
    "-5.8012039600105847814672114227E0", "3.1116436695781989440891606237E-1",
    "-1.52160949662516078556178806805E-1", "2.01365400804030348374776537501E-1",
    "4.47106157277725905176885569043E-2",
]
DP8_BHH_TXT = [
    "0.244094488188976377952755905512E+00", "0.733846688281611857341361741547E+00",
    "0.220588235294117647058823529412E-01",
]
# c12 = c13 = 0.0 is the reference's value (true DOP853 has 1.0): SURVEY Q1
DP8_C_TXT = [
    Z, "0.526001519587677318785587544488E-01", "0.789002279381515978178381316732E-01",
    "0.118350341907227396726757197510E+00", "0.281649658092772603273242802490E+00",
    "0.333333333333333333333333333333E+00", "0.25E+00",
    "0.307692307692307692307692307692E+00", "0.651282051282051282051282051282E+00",
    "0.6E+00", "0.857142857142857142857142857142E+00", Z, Z,
    "0.1E+00", "0.2E+00", "0.777777777777777777777777777778E+00",
]
DP8_D_TXT = [
    ["-0.84289382761090128651353491142E+01", Z, Z, Z, Z,
     "0.56671495351937776962531783590E+00", "-0.30689499459498916912797304727E+01",
     "0.23846676565120698287728149680E+01", "0.21170345824450282767155149946E+01",
     "-0.87139158377797299206789907490E+00", "0.22404374302607882758541771650E+01",
     "0.63157877876946881815570249290E+00", "-0.88990336451333310820698117400E-01",
     "0.18148505520854727256656404962E+02", "-0.91946323924783554000451984436E+01",
     "-0.44360363875948939664310572000E+01"],
    ["0.10427508642579134603413151009E+02", Z, Z, Z, Z,
     "0.24228349177525818288430175319E+03", "0.16520045171727028198505394887E+03",
     "-0.37454675472269020279518312152E+03", "-0.22113666853125306036270938578E+02",
     "0.77334326684722638389603898808E+01", "-0.30674084731089398182061213626E+02",
     "-0.93321305264302278729567221706E+01", "0.15697238121770843886131091075E+02",
     "-0.31139403219565177677282850411E+02", "-0.93529243588444783865713862664E+01",
     "0.35816841486394083752465898540E+02"],
    ["0.19985053242002433820987653617E+02", Z, Z, Z, Z,
     "-0.38703730874935176555105901742E+03", "-0.18917813819516756882830838328E+03",
     "0.52780815920542364900561016686E+03", "-0.11573902539959630126141871134E+02",
     "0.68812326946963000169666922661E+01", "-0.10006050966910838403183860980E+01",
     "0.77771377980534432092869265740E+00", "-0.27782057523535084065932004339E+01",
     "-0.60196695231264120758267380846E+02", "0.84320405506677161018159903784E+02",
     "0.11992291136182789328035130030E+02"],
    ["-0.25693933462703749003312586129E+02", Z, Z, Z, Z,
     "-0.15418974869023643374053993627E+03", "-0.23152937917604549567536039109E+03",
     "0.35763911791061412378285349910E+03", "0.93405324183624310003907691704E+02",
     "-0.37458323136451633156875139351E+02", "0.10409964950896230045147246184E+03",
     "0.29840293426660503123344363579E+02", "-0.43533456590011143754432175058E+02",
     "0.96324553959188282948394950600E+02", "-0.39177261675615439165231486172E+02",
     "-0.14972683625798562581422125276E+03"],
]
DP8_E_TXT = [
    "0.1312004499419488073250102996E-01", Z, Z, Z, Z,
    "-0.1225156446376204440720569753E+01", "-0.4957589496572501915214079952E+00",
    "0.1664377182454986536961530415E+01", "-0.3503288487499736816886487290E+00",
    "0.3341791187130174790297318841E+00", "0.8192320648511571246570742613E-01",
    "-0.2235530786388629525884427845E-01", Z, Z, Z, Z,
]


def hx(v):
    v = float(v)
    if v == 0.0:
        return "0.0"
    return float.hex(v)


def square(rows, n):
    """rows: dict 1-based row -> list; returns n x n dense (0-based) matrix."""
    m = [[0.0] * n for _ in range(n)]
    for i, r in rows.items():
        for j, v in enumerate(r):
            m[i - 1][j] = float(v)
    return m


def c_array_1d(name, vals, comment=""):
    s = "static const double %s[%d] = { /* %s */\n" % (name, len(vals), comment)
    s += "".join("    %s,\n" % hx(v) for v in vals)
    return s + "};\n"


def c_array_2d(name, m, comment=""):
    s = "static const double %s[%d][%d] = { /* %s */\n" % (name, len(m), len(m[0]), comment)
    for r in m:
        s += "    { " + ", ".join(hx(v) for v in r) + " },\n"
    return s + "};\n"


def cx_fn_1d(name, vals):
    body = ", ".join(hx(v) for v in vals)
    return ("    ODE_HD static constexpr double %s(int i) {\n"
            "        constexpr double t[%d] = { %s };\n"
            "        return t[i];\n    }\n" % (name, len(vals), body))


def cx_fn_2d(name, m):
    rows = ",\n            ".join("{ " + ", ".join(hx(v) for v in r) + " }" for r in m)
    return ("    ODE_HD static constexpr double %s(int i, int j) {\n"
            "        constexpr double t[%d][%d] = {\n            %s };\n"
            "        return t[i][j];\n    }\n" % (name, len(m), len(m[0]), rows))


def dev_array_1d(name, vals):
    return "static __constant__ double %s[%d] = { %s };\n" % (name, len(vals), ", ".join(hx(v) for v in vals))


def dev_array_2d(name, m):
    rows = ",\n    ".join("{ " + ", ".join(hx(v) for v in r) + " }" for r in m)
    return "static __constant__ double %s[%d][%d] = {\n    %s };\n" % (name, len(m), len(m[0]), rows)


def main():
    a5 = square(DP5_A, 7)
    a8 = square(DP8_A_TXT, 16)
    b8 = [float(v) for v in DP8_B_TXT]
    bhh8 = [float(v) for v in DP8_BHH_TXT]
    c8 = [float(v) for v in DP8_C_TXT]
    d8 = [[float(v) for v in r] for r in DP8_D_TXT]
    e8 = [float(v) for v in DP8_E_TXT]

    head = ("/* GENERATED by tools/gen_tableau.py -- do not edit.\n"
            " * Dormand-Prince 5(4) and DOP853 coefficients, the values of the reference's\n"
            " * src/butcher_tableau.rs:10-54 and :91-377 (0-based here: X[i-1][j-1] = x(i,j)).\n"
            " * Hex-float literals = the IEEE doubles rustc const-folds from the same expressions. */\n")

    c = head + "#ifndef ODE_TABLEAU_C_H\n#define ODE_TABLEAU_C_H\n\n"
    c += c_array_2d("ODE_DP5_A", a5, "a(i,j), rows 2..7; row 7 = FSAL weights (a72 = 0)")
    c += c_array_1d("ODE_DP5_C", DP5_C, "c(i)")
    c += c_array_1d("ODE_DP5_D", DP5_D, "dense-output d(i), d2 = 0")
    c += c_array_1d("ODE_DP5_E", DP5_E, "error weights e(i), e2 = 0")
    c += c_array_2d("ODE_DP8_A", a8, "a(i,j), rows 2..16; row 13 all zero; rows 14..16 dense stages")
    c += c_array_1d("ODE_DP8_B", b8, "8th-order weights b(i)")
    c += c_array_1d("ODE_DP8_BHH", bhh8, "3rd-order estimator weights")
    c += c_array_1d("ODE_DP8_C", c8, "c(i); c12 = c13 = 0.0 as in the reference (Q1)")
    c += c_array_2d("ODE_DP8_D", d8, "dense d(i,j) for i = 4..7 -> row i-4")
    c += c_array_1d("ODE_DP8_E", e8, "5th-order error weights e(i)")
    c += "\n#endif\n"
    with open(os.path.join(ROOT, "oracle", "ode_tableau_c.h"), "w") as f:
        f.write(c)

    k = head + ("#pragma once\n#ifndef ODE_HD\n#define ODE_HD __host__ __device__ __forceinline__\n#endif\n\n"
                "namespace ode_b200 {\n\n"
                "/* All accessors are 0-based and constexpr so that, with compile-time indices,\n"
                " * coefficients become immediates and structural zeros can be tested with\n"
                " * `if constexpr`. */\n"
                "struct TabDp5 {\n")
    k += cx_fn_2d("a", a5) + cx_fn_1d("c", DP5_C) + cx_fn_1d("d", DP5_D) + cx_fn_1d("e", DP5_E)
    k += "};\n\nstruct TabDp8 {\n"
    k += (cx_fn_2d("a", a8) + cx_fn_1d("b", b8) + cx_fn_1d("bhh", bhh8) + cx_fn_1d("c", c8)
          + cx_fn_2d("d", d8) + cx_fn_1d("e", e8))
    k += "};\n\n"
    k += ("/* The same coefficients in the constant bank (deliberately NOT const, so the compiler cannot\n"
          " * fold them back into immediates): an FP64 instruction reads a constant-bank operand directly\n"
          " * (DMUL R, R, c[0x3][..]), whereas a 64-bit immediate costs two extra UMOV/IMAD.MOV per use --\n"
          " * measured at 26% of all issued instructions before this change (profiles/r1_notes.md).\n"
          " * The constexpr accessors above are still what `if constexpr` tests for structural zeros. */\n"
          "#ifdef __CUDACC__\n")
    k += dev_array_2d("c_dp5_a", a5) + dev_array_1d("c_dp5_c", DP5_C) + dev_array_1d("c_dp5_d", DP5_D)
    k += dev_array_1d("c_dp5_e", DP5_E)
    k += dev_array_2d("c_dp8_a", a8) + dev_array_1d("c_dp8_b", b8) + dev_array_1d("c_dp8_bhh", bhh8)
    k += dev_array_1d("c_dp8_c", c8) + dev_array_2d("c_dp8_d", d8) + dev_array_1d("c_dp8_e", e8)
    k += "#endif\n\n}  // namespace ode_b200\n"
    with open(os.path.join(ROOT, "ode_solvers_b200", "csrc", "ode_tableau.cuh"), "w") as f:
        f.write(k)
    print("wrote oracle/ode_tableau_c.h and ode_solvers_b200/csrc/ode_tableau.cuh")


if __name__ == "__main__":
    main()
