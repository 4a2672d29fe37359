// operators_emul.cpp -- TEST INFRASTRUCTURE.  Runs the operator point functions of chmy.jl_b200/csrc/operators.cuh (the
// same source nvcc compiles into k_box<OperatorF>) on the host, one call of opr_apply per index of the launch box, so that
// tests/test_operators_emulation.py can compare them bit-for-bit with the oracle's independent restatement without a
// GPU.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "../../chmy.jl_b200/csrc/operators.cuh"

extern "C" int operators_emul_run(const OprArgs* g, const int* lo, const int* hi) {
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) opr_apply(*g, i, j, k);
    return 0;
}

// the Float32 instantiation (fields, inv_spacing and every operation in binary32)
extern "C" int operators_emul_run_f32(const OprArgsT<float>* g, const int* lo, const int* hi) {
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) opr_apply(*g, i, j, k);
    return 0;
}

extern "C" int operators_emul_sizeof_args(int f32) { return f32 ? (int)sizeof(OprArgsT<float>) : (int)sizeof(OprArgs); }
