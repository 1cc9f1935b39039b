// Host-side (double precision) instantiation of pde_read_b200/csrc/jet_math.cuh for the CPU
// unit tests: the device code's jet recurrences and their adjoints are checked against
// torch.autograd without a GPU.  Test infrastructure only; not part of the product library.
#include "../../pde_read_b200/csrc/jet_math.cuh"

using namespace pdr;

template <int ACT, int NX, int NT>
static void run_fwd(const double* z, const double* ab, double* y) { act_fwd<ACT, NX, NT, double>(z, ab, y); }

template <int ACT, int NX, int NT>
static void run_vjp(const double* z, const double* ab, const double* yb, double* zb, double* abb) {
    act_vjp<ACT, NX, NT, double>(z, ab, yb, zb, abb);
}

#define FOR_NT(ACT, NX)                                                                     \
    switch (nt) {                                                                            \
        case 0: if (vjp) run_vjp<ACT, NX, 0>(z, ab, yb, out, abb); else run_fwd<ACT, NX, 0>(z, ab, out); return 0; \
        case 1: if (vjp) run_vjp<ACT, NX, 1>(z, ab, yb, out, abb); else run_fwd<ACT, NX, 1>(z, ab, out); return 0; \
        case 2: if (vjp) run_vjp<ACT, NX, 2>(z, ab, yb, out, abb); else run_fwd<ACT, NX, 2>(z, ab, out); return 0; \
        case 3: if (vjp) run_vjp<ACT, NX, 3>(z, ab, yb, out, abb); else run_fwd<ACT, NX, 3>(z, ab, out); return 0; \
        default: return -1;                                                                  \
    }

#define FOR_NX(ACT)                        \
    switch (nx) {                          \
        case 0: FOR_NT(ACT, 0)             \
        case 1: FOR_NT(ACT, 1)             \
        case 2: FOR_NT(ACT, 2)             \
        case 3: FOR_NT(ACT, 3)             \
        case 4: FOR_NT(ACT, 4)             \
        case 5: FOR_NT(ACT, 5)             \
        default: return -1;                \
    }

extern "C" int jet_harness(int act, int nx, int nt, int vjp, const double* z, const double* ab,
                           const double* yb, double* out, double* abb) {
    switch (act) {
        case ACT_TANH: FOR_NX(ACT_TANH)
        case ACT_SIN: FOR_NX(ACT_SIN)
        case ACT_RATIONAL: FOR_NX(ACT_RATIONAL)
        default: return -1;
    }
}
