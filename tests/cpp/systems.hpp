// Test systems with the numbers of the reference fixtures (reference tests/systems.h:11-229), restated
// for the facade tests.  Double integrator under gravity: T = 0.005 s, mass 5 kg, 300 steps.
#pragma once
#include <copra/LMPC.h>
#include <limits>

namespace fixtures {

const double kInf = std::numeric_limits<double>::infinity();

struct QpProblem { // Scilab qld example (tests/systems.h:11-38)
    Eigen::MatrixXd Q, Aeq, Aineq;
    Eigen::VectorXd c, beq, bineq, XL, XU;
    int nrvars = 6, nreqs = 3, nrineqs = 2;
    QpProblem() : Q(6, 6), Aeq(3, 6), Aineq(2, 6), c(6), beq(3), bineq(2), XL(6), XU(6)
    {
        Q = Eigen::MatrixXd::Identity(6, 6);
        c << 1, 2, 3, 4, 5, 6;
        Aeq << 1, -1, 1, 0, 3, 1, -1, 0, -3, -4, 5, 6, 2, 5, 3, 0, 1, 0;
        beq << 1, 2, 3;
        Aineq << 0, 1, 0, 1, 2, -1, -1, 0, 2, 1, 1, 0;
        bineq << -1, 2.5;
        XL << -1000, -10000, 0, -1000, -1000, -1000;
        XU << 10000, 100, 1.5, 100, 100, 1000;
    }
};

struct DoubleIntegrator {
    double T = 0.005, mass = 5;
    int nbStep = 300;
    Eigen::MatrixXd A, B, M, N;
    Eigen::VectorXd c, x0, xd, ud, wx, wu;
    DoubleIntegrator() : A(2, 2), B(2, 1), M(2, 2), N(1, 1), c(2), x0(2), xd(2), ud(1), wx(2), wu(1)
    {
        A << 1, T, 0, 1;
        B << 0.5 * T * T / mass, T / mass;
        c << (-9.81 / 2.) * T * T, -9.81 * T;
        x0 << 0, -5;
        wx << 10, 10000;
        wu << 1e-4;
        M << 1, 0, 0, 1;
        N << 1;
        xd << 0, -1;
        ud << 2;
    }
};

struct BoundedSystem : DoubleIntegrator { // tests/systems.h:42-90
    Eigen::VectorXd uLower, uUpper, xLower, xUpper;
    BoundedSystem() : uLower(1), uUpper(1), xLower(2), xUpper(2)
    {
        uLower.setConstant(-kInf);
        uUpper.setConstant(200);
        xLower.setConstant(-kInf);
        xUpper(0) = kInf;
        xUpper(1) = 0;
    }
};

struct IneqSystem : DoubleIntegrator { // tests/systems.h:94-137
    Eigen::MatrixXd G, E;
    Eigen::VectorXd h, p;
    IneqSystem() : G(1, 1), E(1, 2), h(1), p(1)
    {
        G << 1;
        h << 200;
        E << 0, 1;
        p << 0;
    }
};

struct MixedSystem : DoubleIntegrator { // tests/systems.h:141-181
    Eigen::MatrixXd G, E;
    Eigen::VectorXd p;
    MixedSystem() : G(1, 1), E(1, 2), p(1)
    {
        G << 1;
        E << 0, 1;
        p << 200;
    }
};

struct EqSystem : DoubleIntegrator { // tests/systems.h:186-229
    Eigen::MatrixXd E;
    Eigen::VectorXd p;
    EqSystem() : E(2, 2), p(2)
    {
        x0 << 0, 0;
        xd << 0, 0;
        E.setZero();
        E(0, 0) = 1;
        p = x0;
    }
};

} // namespace fixtures
