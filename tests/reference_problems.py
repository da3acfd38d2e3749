"""The reference's own test problems (active tests of test/basic.jl, test/lpqp.jl, test/2d.jl,
test/3d.jl, test/misc.jl) with their expected optima -- the known-answer fixtures of SURVEY.md 8c.
Each entry: (name, citation, build(model) -> list of variables, expected objective, expected solution or None)."""
import math

from katana_jl_b200 import expr as E

e = math.e
INF = math.inf


def _p(name, cite, build, obj, sol):
    return (name, cite, build, obj, sol)


def poly5(m, x, y):
    m.constraint(x + y, "<=", 5); m.constraint(2 * x - y, "<=", 3); m.constraint(3 * x + 9 * y, ">=", -10)
    m.constraint(10 * x - y, ">=", -20); m.constraint(-x + 2 * y, "<=", 8)


def open6(m, x, y):
    m.constraint(1 * x - 3 * y, "<=", 3); m.constraint(1 * x - 5 * y, "<=", 0); m.constraint(3 * x + 5 * y, ">=", 15)
    m.constraint(7 * x + 2 * y, ">=", 20); m.constraint(9 * x + 1 * y, ">=", 20); m.constraint(3 * x + 7 * y, ">=", 17)


def b_001_01(m):
    x, y = m.variable(), m.variable(); m.objective("Min", x + 0 * y); poly5(m, x, y); return [x, y]
def b_001_02(m):
    x, y = m.variable(), m.variable(); m.objective("Min", (x - 1)**2 + (y - 2)**2); poly5(m, x, y); return [x, y]
def b_002_01(m):
    x, y = m.variable(), m.variable(); m.objective("Min", x + y); open6(m, x, y); return [x, y]
def b_002_02(m):
    x, y = m.variable(), m.variable(); m.objective("Min", (x - 3)**2 + (y - 2)**2); open6(m, x, y); return [x, y]


def disk(sense, obj, bounded=True, extra=None):
    def build(m):
        x, y = (m.variable(-2, 2), m.variable(-2, 2)) if bounded else (m.variable(), m.variable())
        m.objective(sense, obj(x, y)); m.constraint(x**2 + y**2, "<=", 1.0)
        if extra: extra(m, x, y)
        return [x, y]
    return build


def line(m, x, y): m.constraint(x + y, ">=", 1.2)


def parab(sense, obj, extra=None):
    def build(m):
        x, y = m.variable(), m.variable()
        m.objective(sense, obj(x, y)); m.constraint(x**2, "<=", y); m.constraint(-x**2 + 1, ">=", y)
        if extra: extra(m, x, y)
        return [x, y]
    return build


def explog(obj):
    def build(m):
        x, y = m.variable(), m.variable()
        m.objective("Min", obj(x, y))
        m.nlconstraint(E.const(e)**(x - 2.0) - 0.5, "<=", y); m.nlconstraint(E.log(x) + 0.5, ">=", y)
        return [x, y]
    return build


def trig_cap(m):
    """sin / cos rows that ARE convex on the box (the reference's own 106 cases, test/2d.jl:357-401, are commented out upstream as
    non-convex): max y under y <= sin(x), y <= cos(x - 0.5), 0.2 <= x <= 2: both concave there; optimum where the curves cross."""
    x, y = m.variable(0.2, 2.0), m.variable(-1.0, 2.0)
    m.objective("Min", -y + 0 * x)
    m.nlconstraint(E.sin(x), ">=", y); m.nlconstraint(E.cos(x - 0.5), ">=", y)
    return [x, y]


def piecewise_bowl(m):
    """ifelse front-end op: f(x) = ifelse(x <= 1, x^2, 2x - 1) is convex and C1; min y - 1.5 x under f(x) <= y: f'(x) = 1.5 at x = 0.75."""
    x, y = m.variable(-2.0, 3.0), m.variable(-5.0, 10.0)
    m.objective("Min", y - 1.5 * x)
    m.nlconstraint(E.ifelse(E.le(x, 1.0), x**2, 2.0 * x - 1.0), "<=", y)
    return [x, y]


def b_108_01(m):
    x, y = m.variable(0, INF), m.variable(0, INF)
    m.objective("Min", (x - 1.0)**2 + (y - 0.75)**2)
    m.nlconstraint(2 * x**2 - 4 * x * y - 4 * x + 4, "<=", y); m.constraint(y**2, "<=", -x + 2)
    return [x, y]


def nlobj_disk(obj):
    def build(m):
        x, y = m.variable(), m.variable()
        m.nlobjective("Min", obj(x, y)); m.constraint(x**2 + y**2, "<=", 1.0)
        return [x, y]
    return build


def sphere3(obj):
    def build(m):
        x, y, z = m.variable(), m.variable(), m.variable()
        m.objective("Min", obj(x, y, z)); m.constraint(x**2 + y**2 + z**2, "<=", 1.0)
        return [x, y, z]
    return build


def parab3(obj):
    def build(m):
        x, y, z = m.variable(), m.variable(), m.variable()
        m.objective("Min", obj(x, y, z)); m.constraint(x**2 + y**2, "<=", z); m.constraint(x**2 + y**2, "<=", -z + 1)
        return [x, y, z]
    return build


def b_203_01(m):
    x, y, z = m.variable(), m.variable(), m.variable()
    m.objective("Min", x + y); m.nlconstraint(E.sqrt(x**2 + y**2), "<=", z - 0.25); m.constraint(x**2 + y**2, "<=", -z + 1)
    return [x, y, z]


def b_205_01(m):
    x, y, z = m.variable(), m.variable(0, INF), m.variable()
    m.objective("Max", y + 0 * x)
    m.nlconstraint(y * E.const(e)**(x / y), "<=", z); m.nlconstraint(y * E.const(e)**((-x) / y), "<=", z)
    m.constraint(x**2 + y**2, "<=", -z + 5)
    return [x, y, z]


def nl_sphere3(c):
    def build(m):
        x, y, z = m.variable(), m.variable(), m.variable()
        m.objective("Min", (x - c)**2 + (y - c)**2 + (z - c)**2); m.nlconstraint(x**2 + y**2 + z**2, "<=", 1.0)
        return [x, y, z]
    return build


def nd_sphere(n, norm_form):
    def build(m):
        v = m.variables(n)
        m.objective("Min", E.sum_([-x for x in v]))
        s = E.sum_([x**2 for x in v])
        m.nlconstraint(E.sqrt(s) if norm_form else s, "<=", 1.0)
        return v
    return build


r2, r3 = math.sqrt(2), math.sqrt(3)
PROBLEMS = [
    _p("001_01", "test/lpqp.jl:7-27", b_001_01, -2.0430107680954848, [-2.0430107680954848, -0.4301075068564087]),
    _p("001_02", "test/lpqp.jl:30-50 / test/basic.jl:80-102", b_001_02, 0.0, [1.0, 2.0]),
    _p("002_01", "test/lpqp.jl:53-73", b_002_01, 3.9655172067026196, [2.4137930845761546, 1.5517241221264648]),
    _p("002_02", "test/lpqp.jl:76-97", b_002_02, 0.0, [3.0, 2.0]),
    _p("101_01", "test/2d.jl:5-20", disk("Min", lambda x, y: -x - y), -2 / r2, [1 / r2, 1 / r2]),
    _p("101_02", "test/2d.jl:23-38", disk("Min", lambda x, y: -x + 0 * y), -1.0, [1.0, 0.0]),
    _p("101_03", "test/2d.jl:41-56", disk("Max", lambda x, y: x + 0 * y), 1.0, [1.0, 0.0]),
    _p("102_01", "test/2d.jl:60-76", disk("Min", lambda x, y: -x + 0 * y, False, line), -0.974165743715913, [0.974165743715913, 0.2258342542139504]),
    _p("102_02", "test/2d.jl:79-96", disk("Min", lambda x, y: x + y, False, line), 1.2, None),
    _p("102_03", "test/2d.jl:99-115", disk("Max", lambda x, y: x + y, False, line), 2 / r2, [1 / r2, 1 / r2]),
    _p("102_04", "test/2d.jl:118-134", disk("Min", lambda x, y: x**2 + y**2, False, line), 0.72, [0.6, 0.6]),
    _p("102_05", "test/2d.jl:137-153", disk("Min", lambda x, y: (x - 0.65)**2 + (y - 0.65)**2, False, line), 0.0, [0.65, 0.65]),
    _p("103_01", "test/2d.jl:157-173", parab("Min", lambda x, y: y + 0 * x), 0.0, [0.0, 0.0]),
    _p("103_02", "test/2d.jl:176-192", parab("Min", lambda x, y: -y + 0 * x), -1.0, [0.0, 1.0]),
    _p("103_03", "test/2d.jl:195-211", parab("Min", lambda x, y: -x - y), -5 / 4, [2 / 4, 3 / 4]),
    _p("103_04", "test/2d.jl:214-230", parab("Min", lambda x, y: x + y), -1 / 4, [-2 / 4, 1 / 4]),
    _p("103_05", "test/2d.jl:233-249", parab("Min", lambda x, y: -x + 0 * y), -1 / r2, [1 / r2, 1 / 2]),
    _p("104_01", "test/2d.jl:253-271", parab("Min", lambda x, y: -x + 0 * y, lambda m, x, y: m.constraint(x**2 + (y - 0.5)**2, "<=", 1.0)), -1 / r2, [1 / r2, 1 / 2]),
    _p("105_01", "test/2d.jl:275-291", explog(lambda x, y: -x - y), -4.176004405036646, [2.687422019398147, 1.488582385638499]),
    _p("105_04", "test/2d.jl:338-354", explog(lambda x, y: -x + y), -3 / 2, [2.0, 1 / 2]),
    _p("107_01", "test/2d.jl:405-420", disk("Min", lambda x, y: (x - 0.5)**2 + (y - 0.5)**2, False), 0.0, [0.5, 0.5]),
    _p("107_02", "test/2d.jl:423-438", disk("Min", lambda x, y: (x - 1.0)**2 + (y - 1.0)**2, False), 0.17157287363083387, [1 / r2, 1 / r2]),
    _p("106_xx", "sin/cos front-end ops; convex variant of test/2d.jl:357-401", trig_cap, -0.8600655610487502, [1.0353981633974483, 0.8600655610487502]),
    _p("ifelse_xx", "ifelse / comparison front-end ops (SURVEY 8f item 4)", piecewise_bowl, -0.5625, [0.75, 0.5625]),
    _p("108_01", "test/2d.jl:460-476", b_108_01, 0.0, [1.0, 0.75]),
    _p("110_01", "test/2d.jl:603-618", nlobj_disk(lambda x, y: E.const(e)**x), e**-1, [-1.0, 0.0]),
    _p("110_02", "test/2d.jl:621-636", nlobj_disk(lambda x, y: E.const(e)**x + E.const(e)**y), 2 * e**(-1 / r2), [-1 / r2, -1 / r2]),
    _p("110_03", "test/2d.jl:639-654", nlobj_disk(lambda x, y: E.const(e)**(x + y)), e**(-2 / r2), [-1 / r2, -1 / r2]),
    _p("201_01", "test/3d.jl:5-22", sphere3(lambda x, y, z: -(x + y + z)), -3 / r3, [1 / r3] * 3),
    _p("201_02", "test/3d.jl:25-42", sphere3(lambda x, y, z: -x + 0 * y + 0 * z), -1.0, [1.0, 0.0, 0.0]),
    _p("202_01", "test/3d.jl:46-63", parab3(lambda x, y, z: -z + 0 * x), -1.0, [0.0, 0.0, 1.0]),
    _p("202_02", "test/3d.jl:66-84", parab3(lambda x, y, z: z + 0 * x), 0.0, [0.0, 0.0, 0.0]),
    _p("202_03", "test/3d.jl:87-105", parab3(lambda x, y, z: -(x + y + 2 * z)), -9 / 4, [1 / 4, 1 / 4, 7 / 8]),
    _p("202_04", "test/3d.jl:108-127", parab3(lambda x, y, z: x + y + 2 * z), -1 / 4, [-1 / 4, -1 / 4, 1 / 8]),
    _p("202_05", "test/3d.jl:130-148", parab3(lambda x, y, z: x + y + 0 * z), -1.0, [-1 / 2, -1 / 2, 1 / 2]),
    _p("203_01", "test/3d.jl:153-171", b_203_01, -1 / r2, [-math.sqrt(1 / 8), -math.sqrt(1 / 8), 3 / 4]),
    _p("205_01", "test/3d.jl:221-240", b_205_01, 1.7912878443121907, [0.0, 1.7912878443121907, 1.7912878443121907]),
    _p("210_01", "test/3d.jl:271-287", nl_sphere3(0.5), 0.0, [0.5] * 3),
    _p("210_02", "test/3d.jl:290-307", nl_sphere3(1.0), 0.535898380052066, [1 / r3] * 3),
] + [_p(f"501_01_n{n}", "test/misc.jl:4-30", nd_sphere(n, False), -n / math.sqrt(n), [1 / math.sqrt(n)] * n) for n in (1, 2, 3, 5, 8, 13)] \
  + [_p(f"501_02_n{n}", "test/misc.jl:33-57", nd_sphere(n, True), -n / math.sqrt(n), [1 / math.sqrt(n)] * n) for n in (2, 3, 5, 8, 13)]   # the reference runs n = 1..20; n = 20 needs minutes of host LP re-solves with the scipy stand-in
