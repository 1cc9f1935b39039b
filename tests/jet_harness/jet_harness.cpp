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

// "row" forms used by the reverse kernels for tanh: slot 0 of the row carries y_0 = tanh(z_0); the VJP takes
// the forward jet y as an input.  z_0 itself is poisoned to prove it is never read.
template <int NX, int NT>
static void run_row_fwd(const double* z, double* y) {
    constexpr int J = 1 + NX + NT;
    double row[J], w[J];
    for (int j = 0; j < J; ++j) row[j] = z[j];
    row[0] = std::tanh(z[0]);
    tanh_series<NX, NT, double>(row[0], row, y, w);
}
template <int NX, int NT>
static void run_row_vjp(const double* z, const double* yb, double* zb) {
    constexpr int J = 1 + NX + NT;
    double row[J], y[J], w[J];
    tanh_fwd<NX, NT, double>(z, y, w);
    for (int j = 0; j < J; ++j) row[j] = z[j];
    row[0] = std::nan("");
    tanh_vjp_y<NX, NT, double>(row, y, yb, zb);
}

template <int NX, int NT>
static void run_rat_vjp_y(const double* z, const double* ab, const double* yb, double* zb, double* abb) {
    constexpr int J = 1 + NX + NT;
    double y[J];
    rational_fwd<NX, NT, double>(z, ab, y);
    rational_vjp_y<NX, NT, double>(z, y, ab, yb, zb, abb);
}
#define RAT_NT(NX)                                                             \
    switch (nt) {                                                               \
        case 0: run_rat_vjp_y<NX, 0>(z, ab, yb, out, abb); return 0;            \
        case 1: run_rat_vjp_y<NX, 1>(z, ab, yb, out, abb); return 0;            \
        case 2: run_rat_vjp_y<NX, 2>(z, ab, yb, out, abb); return 0;            \
        case 3: run_rat_vjp_y<NX, 3>(z, ab, yb, out, abb); return 0;            \
        default: return -1;                                                     \
    }
extern "C" int jet_harness_rational_vjp_y(int nx, int nt, const double* z, const double* ab, const double* yb,
                                          double* out, double* abb) {
    switch (nx) {
        case 0: RAT_NT(0)
        case 1: RAT_NT(1)
        case 2: RAT_NT(2)
        case 3: RAT_NT(3)
        case 4: RAT_NT(4)
        case 5: RAT_NT(5)
        default: return -1;
    }
}

#define ROW_NT(NX)                                                                          \
    switch (nt) {                                                                            \
        case 0: if (vjp) run_row_vjp<NX, 0>(z, yb, out); else run_row_fwd<NX, 0>(z, out); return 0; \
        case 1: if (vjp) run_row_vjp<NX, 1>(z, yb, out); else run_row_fwd<NX, 1>(z, out); return 0; \
        case 2: if (vjp) run_row_vjp<NX, 2>(z, yb, out); else run_row_fwd<NX, 2>(z, out); return 0; \
        case 3: if (vjp) run_row_vjp<NX, 3>(z, yb, out); else run_row_fwd<NX, 3>(z, out); return 0; \
        default: return -1;                                                                  \
    }

extern "C" int jet_harness_tanh_row(int nx, int nt, int vjp, const double* z, const double* yb, double* out) {
    switch (nx) {
        case 0: ROW_NT(0)
        case 1: ROW_NT(1)
        case 2: ROW_NT(2)
        case 3: ROW_NT(3)
        case 4: ROW_NT(4)
        case 5: ROW_NT(5)
        default: return -1;
    }
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
