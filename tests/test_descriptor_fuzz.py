"""
CPU-only robustness of the C ABI's argument checks: a binding in another language hands the library plain structs, so a
malformed descriptor must come back as an error code -- never as a crash, a hang or an out-of-bounds read.

A child process (a crash would otherwise take pytest down with it) builds valid launch descriptors for every solver op in
2D and 3D on descriptor-only fields, then mutates them thousands of times -- op ids, dimensionality, sizes, counts, flags,
batch kinds and counts, outer widths, operator ids, field pointers swapped among valid fields of other shapes / element
types or set to NULL -- and calls chmy_validate_launch and chmy_launch_split_plan on each.  Field pointers are only ever
NULL or valid handles: a C library cannot vet an arbitrary address.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import ctypes as C, random, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(tests)r)
import chmy_b200 as ch
from chmy_b200 import _lib as L
from test_launch_validation import _NoArch, grid, F, vec, ten
lib = L.lib()
rng = random.Random(int(sys.argv[1]))
descs, pool, keep = [], [], []
for n in ((12, 10), (12, 10, 8)):
    g = grid(ch, n); nd = len(n)
    la = ch.Launcher(_NoArch(), g, outer_width=(4,) * nd)
    V, rV, qT, tau, tau_old = vec(ch, g), vec(ch, g), vec(ch, g), ten(ch, g), ten(ch, g)
    Pr, dV, T, To = F(ch, g), F(ch, g), F(ch, g), F(ch, g)
    rho_loc = tuple(ch.Vertex() if i == nd - 1 else ch.Center() for i in range(nd))
    rho = F(ch, g, rho_loc)
    bcV = [(c, {a: (ch.Dirichlet() if a == "xyz"[i] else ch.Neumann()) for a in "xyz"[:nd]}) for i, c in enumerate(V)]
    vfld = F(ch, ch.UniformGrid(_NoArch(), origin=(0.0,) * (nd - 1), extent=(1.0,) * (nd - 1), dims=n[1:]))
    launches = [
        ((ch.update_old_, (T, tau, To, tau_old)), None),
        ((ch.update_stress_, (tau, Pr, dV, V, tau_old, 10.0, 0.1, 1.0, 0.07, 0.3, 0.2, g)), None),
        ((ch.update_velocity_, (V, rV, Pr, tau, rho, 0.1, 0.01, g)), ch.batch(g, *bcV)),
        ((ch.update_thermal_flux_, (qT, T, V, 1e-4, g)), None),
        ((ch.update_thermal_, (T, To, qT, 0.07, g)), ch.batch(g, (T, {"x": ch.Dirichlet(vfld)}))),
    ]
    if nd == 2:
        q, Cf = vec(ch, g), F(ch, g)
        launches += [((ch.compute_q_, (q, Cf, 1.0, g)), None), ((ch.update_C_, (Cf, q, 0.01, g)), ch.batch(g, (Cf, ch.Neumann(2.0))))]
    for oa, bc in launches:
        d = la.describe(None, g, oa, bc=bc)
        L.check(lib.chmy_validate_launch(C.byref(d)))
        descs.append(d)
    g32 = grid(ch, n, np.float32)
    others = [F(ch, g32), F(ch, grid(ch, tuple(x + 1 for x in n))), vfld]
    keep += [V, rV, qT, tau, tau_old, Pr, dV, T, To, rho, others, launches]
    for f in list(V) + list(tau) + [Pr, T, rho] + others:
        pool.append(f.handle.value)
INTS = [-(1 << 31), -7, -1, 0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 63, 64, 255, 1 << 20, (1 << 31) - 1]
BIG = INTS + [1 << 40, -(1 << 40), (1 << 62)]
def ptr():
    return None if rng.random() < 0.3 else rng.choice(pool)
def mutate(d):
    k = rng.randrange(16)
    if k == 0: d.op = rng.choice(INTS)
    elif k == 1: d.grid.ndims = rng.choice(INTS)
    elif k == 2: d.grid.n[rng.randrange(3)] = rng.choice(BIG)
    elif k == 3: d.nfields = rng.choice(INTS)
    elif k == 4: d.nscalars = rng.choice(INTS)
    elif k == 5: d.flags = rng.choice(INTS)
    elif k == 6: d.fields[rng.randrange(len(d.fields))] = ptr()
    elif k == 7: d.has_bc = rng.choice(INTS)
    elif k == 8: d.has_outer_width = rng.choice(INTS)
    elif k == 9: d.outer_width[rng.randrange(3)] = rng.choice(BIG)
    elif k == 10: d.bc[rng.randrange(3)][rng.randrange(2)].kind = rng.choice(INTS)
    elif k == 11: d.bc[rng.randrange(3)][rng.randrange(2)].nfields = rng.choice(INTS)
    elif k == 12:
        b = d.bc[rng.randrange(3)][rng.randrange(2)]
        q = rng.randrange(len(b.fields))
        b.fields[q] = ptr()
        b.bc_kind[q] = rng.choice(INTS)
        b.value_field[q] = ptr() if rng.random() < 0.5 else None
    elif k == 13: d.oper, d.oper_dim = rng.choice(INTS), rng.choice(INTS)
    elif k == 14: d.grid.connectivity[rng.randrange(3)][rng.randrange(2)] = rng.choice(INTS)
    else:
        x = rng.choice([float("nan"), float("inf"), 0.0, -0.0, -1.0, 1e308])
        d.grid.spacing[rng.randrange(3)] = x; d.grid.inv_spacing[rng.randrange(3)] = x; d.scalars[rng.randrange(len(d.scalars))] = x
ok = err = 0
split, wl, wr = C.c_int32(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
for it in range(int(sys.argv[2])):
    d = L.LaunchDesc.from_buffer_copy(bytes(rng.choice(descs)))
    for _ in range(rng.randrange(1, 4)):
        mutate(d)
    rc = lib.chmy_validate_launch(C.byref(d))
    assert isinstance(rc, int) and -5 <= rc <= 0, rc
    ok += rc == 0; err += rc != 0
    pref = None if rng.random() < 0.5 else (C.c_int32 * 3)(*[rng.choice(INTS) for _ in range(3)])
    rc2 = lib.chmy_launch_split_plan(C.byref(d), pref, rng.randrange(2), C.byref(split), wl, wr)
    assert isinstance(rc2, int) and -5 <= rc2 <= 0, rc2
    if rc2 == 0 and split.value:
        assert all(0 <= w < (1 << 30) for w in list(wl) + list(wr))
print("FUZZ-DONE", ok, err)
'''


def test_mutated_descriptors_are_refused_not_crashed_on():
    code = CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests")}
    total_ok = total_err = 0
    for seed in (1, 2, 3):
        r = subprocess.run([sys.executable, "-c", code, str(seed), "4000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0 and "FUZZ-DONE" in r.stdout, (r.returncode, r.stdout[-1500:], r.stderr[-3000:])
        _, ok, err = r.stdout.strip().splitlines()[-1].split()
        total_ok += int(ok)
        total_err += int(err)
    assert total_err > 3000 and total_ok > 100, (total_ok, total_err)      # both outcomes are exercised
