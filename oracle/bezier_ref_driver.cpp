// oracle/bezier_ref_driver.cpp -- C entry point around the REFERENCE's own Bernstein evaluators
// (global_planner/include/global_planner/utils/bezier_base.h:77-115 getPos / getVel / getAcc, constructor and tables in
// global_planner/src/utils/bezier_base.cpp, both compiled unmodified from /root/reference against oracle/shim).
// TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/libbezier_ref.so, which pins the numpy restatement
// oracle_py.bezier_sample and produces tests/golden/bezier_ref.npz (tests/golden/make_bezier_golden.py); the GPU kernel
// bezier_sample_kernel is compared with that fixture.  Scaling as the reference's callers apply it:
// position = time * getPos (teach_repeat_planner.cpp:1557-1560), velocity = getVel, acceleration = getAcc / time (:681-682).
#include <global_planner/utils/bezier_base.h>

extern "C" int bezier_ref_sample(int B, int N, int S, const double *bez /*[B][N][18]*/, const double *times /*[B][N]*/,
                                 double *pos, double *vel, double *acc /*[B][N][S][3] each*/) {
    Bernstein bern(3.0);          // teach_repeat_planner.cpp:1176 constructs it with the minimise order
    bern.setFixedOrder(5);        // poly_order 5 (global_planner.launch), ctrl_num1D = 6
    for (int b = 0; b < B; b++) {
        Eigen::MatrixXd coeff = Eigen::MatrixXd::Zero(N, 18);
        for (int k = 0; k < N; k++)
            for (int c = 0; c < 18; c++) coeff(k, c) = bez[((size_t)b * N + k) * 18 + c];
        for (int k = 0; k < N; k++) {
            const double T = times[(size_t)b * N + k];
            for (int j = 0; j < S; j++) {
                const double s = S > 1 ? (double)j / (double)(S - 1) : 0.0;
                const Eigen::Vector3d p = bern.getPos(coeff, k, s), v = bern.getVel(coeff, k, s), a = bern.getAcc(coeff, k, s);
                const size_t o = (((size_t)b * N + k) * S + j) * 3;
                for (int d = 0; d < 3; d++) { pos[o + d] = T * p(d); vel[o + d] = v(d); acc[o + d] = a(d) / T; }
            }
        }
    }
    return 0;
}
