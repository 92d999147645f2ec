// Shared device-side definitions: RODAS4 tableau, step-size controller, kernel argument block.
//
// Integrator: RODAS4 (Hairer & Wanner, "Solving ODEs II", sec. IV.7/IV.10): 6-stage stiffly accurate
// Rosenbrock method of order 4 with an embedded order-3 solution; gamma = 1/4; L-stable.  In the
// implementation form each stage solves (I/(h*gamma) - J) U_i = f(y + sum a_ij U_j) + sum (c_ij/h) U_j,
// and y_{n+1} = y_n + sum_{j<=4} a_5j U_j + U_5 + U_6 with error estimate U_6.  All models on this
// path are autonomous, so the time-derivative terms of the method vanish.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pk {

namespace rodas4 {
constexpr double GAMMA = 0.25;
constexpr double A21 = 0.1544000000000000e+01;
constexpr double A31 = 0.9466785280815826e+00, A32 = 0.2557011698983284e+00;
constexpr double A41 = 0.3314825187068521e+01, A42 = 0.2896124015972201e+01, A43 = 0.9986419139977817e+00;
constexpr double A51 = 0.1221224509226641e+01, A52 = 0.6019134481288629e+01, A53 = 0.1253708332932087e+02,
                 A54 = -0.6878860361058950e+00;
constexpr double C21 = -0.5668800000000000e+01;
constexpr double C31 = -0.2430093356833875e+01, C32 = -0.2063599157091915e+00;
constexpr double C41 = -0.1073529058151375e+00, C42 = -0.9594562251023355e+01, C43 = -0.2047028614809616e+02;
constexpr double C51 = 0.7496443313967647e+01, C52 = -0.1024680431464352e+02, C53 = -0.3399990352819905e+02,
                 C54 = 0.1170890893206160e+02;
constexpr double C61 = 0.8083246795921522e+01, C62 = -0.7981132988064893e+01, C63 = -0.3152159432874371e+02,
                 C64 = 0.1631930543123136e+02, C65 = -0.6058818238834054e+01;
}  // namespace rodas4

// Step-size control (per system): elementary controller with safety 0.9, growth in [1/5, 6] per
// step, plus Gustafsson's predictive correction after the second accepted step (as in RODAS).
struct StepCtl {
    double h;        // proposal for the next step
    double hacc;     // last accepted step
    double erracc;   // its error
    int naccpt;
    int rejected_last;
};

constexpr double CTL_SAFE = 0.9;
constexpr double CTL_FAC_SHRINK = 5.0;       // h_new >= h / 5
constexpr double CTL_FAC_GROW = 1.0 / 6.0;   // h_new <= 6 h
constexpr double LAND_STRETCH = 1.03;        // stretch the step by <=3% to land on an output time

__device__ __forceinline__ double ctl_factor(double err) {
    // err^(1/4)/safe, clamped.  err == 0 -> maximal growth.
    double f = sqrt(sqrt(err)) * (1.0 / CTL_SAFE);
    return fmax(CTL_FAC_GROW, fmin(CTL_FAC_SHRINK, f));
}

// On acceptance of a step of size hh with error err: returns the next proposal.
__device__ __forceinline__ double ctl_accept(StepCtl& c, double hh, double err) {
    double fac = ctl_factor(err);
    if (c.naccpt > 0) {
        double r = err * err / c.erracc;
        double fg = (c.hacc / hh) * sqrt(sqrt(r)) * (1.0 / CTL_SAFE);
        fg = fmax(CTL_FAC_GROW, fmin(CTL_FAC_SHRINK, fg));
        fac = fmax(fac, fg);
    }
    c.hacc = hh;
    c.erracc = fmax(1.0e-2, err);
    c.naccpt++;
    double hnew = hh / fac;
    if (c.rejected_last) hnew = fmin(hnew, hh);
    c.rejected_last = 0;
    return hnew;
}

__device__ __forceinline__ double ctl_reject(StepCtl& c, double hh, double err) {
    c.rejected_last = 1;
    return hh / ctl_factor(err);
}

// Argument block of the local-model kernels (device pointers only).
struct LocalArgs {
    long long B;
    int T, ns, n, P, L;
    const double* params;   // [B,P]
    const double* y0;       // [n] or [B, y0_stride]
    long long y0_stride;
    const double* t;        // [T]
    double rtol, atol;
    int max_steps, normalize, log_params, y_metric;
    double* out_sol;
    double* out_flat;
    double* out_Y;
    double* out_ssr;
    double* out_score;
    int* out_status;
    int* out_nsteps;
    int* out_nrej;
    const double* target;
    const double* sigma;
    const int* group;
    int sigma_len;
    double lam;
    double w_alpha, w_beta, w_gamma, w_delta, w_mu;
    unsigned long long* counter;   // work queue head
};

constexpr int RNA_OFFSET = 5;   // models/distmod.py:125  sol[5:, 0]

// Per-system accumulators of the fused outputs (per lane; the warp kernel reduces them at the end).
struct EpiAcc {
    double ssr, sr, sr2;     // weighted SSR; score_fit sums of r and r^2
    double s1, s2, dyn;      // Morris Y sums
};

}  // namespace pk
