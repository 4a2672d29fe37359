// bc_emul.cpp -- TEST INFRASTRUCTURE.  Runs the per-point bodies of the boundary-condition and halo-slab kernels
// (chmy.jl_b200/csrc/bc_kernels.cuh: the same source nvcc compiles into k_bc_dim and k_slab) on the host, one call per
// thread of the launch grid bc.cu would start (including the threads past the edge that the kernels mask off), so that
// tests/test_bc_emulation.py can compare them bit-for-bit with the oracle without a GPU.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "../../chmy.jl_b200/csrc/bc_kernels.cuh"

// grid of run_bc_dim (bc.cu): x = ceil(nt0 / 128) blocks of 128 threads, y = nt1
template <class T>
static int run_bc(const BcBatchDev<T>* b) {
    const int nx = (b->nt[0] + 127) / 128 * 128;
    for (int c = 0; c < b->nt[1]; ++c)
        for (int a = 0; a < nx; ++a) {
            if (a >= b->nt[0]) continue;        // the kernel's own guard
            bc_point(*b, a, c);
        }
    return 0;
}
extern "C" int bc_emul_run(const BcBatchDev<double>* b) { return run_bc(b); }
extern "C" int bc_emul_run_f32(const BcBatchDev<float>* b) { return run_bc(b); }

// grid of run_slab (bc.cu): blocks of 64 x (4 | 1) threads over the largest slab of the batch, z = field
template <class T>
static int run_slab(const SlabBatch<T>* b, T* buf, int pack) {
    int m0 = 1, m1 = 1;
    for (int q = 0; q < b->n; ++q) { m0 = b->e[q].e0 > m0 ? b->e[q].e0 : m0; m1 = b->e[q].e1 > m1 ? b->e[q].e1 : m1; }
    const int by = m1 > 1 ? 4 : 1;
    const int nx = (m0 + 63) / 64 * 64, ny = (m1 + by - 1) / by * by;
    for (int q = 0; q < b->n; ++q)
        for (int c = 0; c < ny; ++c)
            for (int a = 0; a < nx; ++a) {
                if (pack) slab_point<true, T>(*b, buf, q, a, c);
                else slab_point<false, T>(*b, buf, q, a, c);
            }
    return 0;
}
extern "C" int slab_emul_run(const SlabBatch<double>* b, double* buf, int pack) { return run_slab(b, buf, pack); }
extern "C" int slab_emul_run_f32(const SlabBatch<float>* b, float* buf, int pack) { return run_slab(b, buf, pack); }

// grid of run_bc_all_t (bc.cu): x = ceil(m0 / 128) blocks of 128 threads, y = m1 (largest face extents over the dims),
// z = (field, dim, side).  `rev`: visit the threads in reverse order -- the result may not depend on the order, because no
// thread reads a cell another thread writes.
template <class T>
static int run_bc_all(const BcAllDev<T>* b, int rev) {
    int m0 = 1, m1 = 1;
    for (int D = 0; D < b->nd; ++D) {
        int e[2] = {1, 1}, t = 0;
        for (int a = 0; a < b->nd; ++a)
            if (a != D) e[t++] = b->n[a] + 3;
        m0 = e[0] > m0 ? e[0] : m0; m1 = e[1] > m1 ? e[1] : m1;
    }
    const int nx = (m0 + 127) / 128 * 128, nz = b->nact;
    const long long total = (long long)nz * m1 * nx;
    for (long long i = 0; i < total; ++i) {
        const long long t = rev ? total - 1 - i : i;
        const int a = (int)(t % nx), c = (int)((t / nx) % m1), z = b->act[(int)(t / ((long long)nx * m1))];
        const int s = z & 1, D = (z >> 1) % 3, q = z / 6;
        int nt[2] = {1, 1}, k = 0;
        for (int d = 0; d < b->nd; ++d)
            if (d != D) nt[k++] = b->n[d] + 3;
        if (a >= nt[0] || c >= nt[1]) continue;      // the kernel's own guards
        bc_all_point(*b, q, D, s, a, c);
    }
    return 0;
}
extern "C" int bc_all_emul_run(const BcAllDev<double>* b, int rev) { return run_bc_all(b, rev); }
extern "C" int bc_all_emul_run_f32(const BcAllDev<float>* b, int rev) { return run_bc_all(b, rev); }

extern "C" int bc_emul_sizeof(int what, int f32) {
    if (what == 0) return f32 ? (int)sizeof(BcBatchDev<float>) : (int)sizeof(BcBatchDev<double>);
    if (what == 2) return f32 ? (int)sizeof(BcAllDev<float>) : (int)sizeof(BcAllDev<double>);
    return f32 ? (int)sizeof(SlabBatch<float>) : (int)sizeof(SlabBatch<double>);
}
