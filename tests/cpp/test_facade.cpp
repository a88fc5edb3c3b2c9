// test_facade.cpp -- the reference's own test scenarios (tests/TestLMPC.cpp, TestSolvers.cpp,
// TestLMPC_InitialState.cpp) re-expressed against the B200 facade.  Same fixtures, same assertions
// (properties, tolerances, exception types).  `./test_facade cpu` runs the groups that need no GPU
// (shape logic, error handling, autoSpan); `./test_facade gpu` runs everything.
#include "systems.hpp"

#include <copra/InitialStateLMPC.h>
#include <copra/LMPC.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

static int g_fail = 0, g_checks = 0;
#define REQUIRE(cond)                                                                  \
    do {                                                                               \
        ++g_checks;                                                                    \
        if (!(cond)) { ++g_fail; std::printf("  FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define REQUIRE_LE(a, b) REQUIRE((a) <= (b))
#define REQUIRE_THROWS_AS(expr, type)                                                   \
    do {                                                                                \
        ++g_checks;                                                                     \
        bool ok_ = false;                                                               \
        try { expr; } catch (const type&) { ok_ = true; } catch (...) {}                \
        if (!ok_) { ++g_fail; std::printf("  FAILED %s:%d  %s should throw %s\n", __FILE__, __LINE__, #expr, #type); } \
    } while (0)
#define REQUIRE_NOTHROW(expr)                                                           \
    do {                                                                                \
        ++g_checks;                                                                     \
        try { expr; } catch (const std::exception& e_) { ++g_fail; std::printf("  FAILED %s:%d  %s threw %s\n", __FILE__, __LINE__, #expr, e_.what()); } \
    } while (0)

using namespace fixtures;

static Eigen::VectorXd spanVec(const Eigen::VectorXd& v, int n)
{
    Eigen::VectorXd out(v.rows() * n);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < v.rows(); ++k) out(i * v.rows() + k) = v(k);
    return out;
}
static Eigen::MatrixXd spanMat(const Eigen::MatrixXd& m, int n, int addCols = 0)
{
    Eigen::MatrixXd out = Eigen::MatrixXd::Zero(m.rows() * n, m.cols() * (n + addCols));
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < m.cols(); ++c)
            for (int r = 0; r < m.rows(); ++r) out(i * m.rows() + r, i * m.cols() + c) = m(r, c);
    return out;
}
static void splitTraj(const Eigen::VectorXd& full, Eigen::VectorXd& pos, Eigen::VectorXd& vel)
{
    const int len = int(full.rows() / 2);
    pos.resize(len); vel.resize(len);
    for (int i = 0; i < len; ++i) { pos(i) = full(2 * i); vel(i) = full(2 * i + 1); }
}

// ------------------------------------------------------------------------------------------------
// groups that need no GPU
// ------------------------------------------------------------------------------------------------
static void errorHandlers() // tests/TestLMPC.cpp:949-1087
{
    IneqSystem s;
    {
        auto ps = std::make_shared<copra::PreviewSystem>();
        REQUIRE_THROWS_AS(ps->system(Eigen::MatrixXd::Ones(5, 2), s.B, s.c, s.x0, s.nbStep), std::domain_error);
        REQUIRE_THROWS_AS(ps->system(Eigen::MatrixXd::Ones(2, 5), s.B, s.c, s.x0, s.nbStep), std::domain_error);
        REQUIRE_THROWS_AS(ps->system(s.A, Eigen::MatrixXd::Ones(5, 1), s.c, s.x0, s.nbStep), std::domain_error);
        REQUIRE_THROWS_AS(ps->system(s.A, s.B, Eigen::VectorXd::Ones(5), s.x0, s.nbStep), std::domain_error);
        REQUIRE_THROWS_AS(ps->system(s.A, s.B, s.c, s.x0, -1), std::domain_error);
    }
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    { // weights
        auto cost = std::make_shared<copra::TrajectoryCost>(s.M, s.xd);
        REQUIRE_NOTHROW(cost->weight(2));
        REQUIRE_THROWS_AS(cost->weights(Eigen::VectorXd::Ones(5)), std::domain_error);
        REQUIRE_NOTHROW(cost->weights(s.wx));
        REQUIRE_NOTHROW(controller.addCost(cost));
        REQUIRE_NOTHROW(cost->weights(Eigen::VectorXd::Ones(2)));
    }
    { // costs
        auto badM = std::make_shared<copra::TrajectoryCost>(Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(2));
        REQUIRE_THROWS_AS(controller.addCost(badM), std::domain_error);
        auto badP = std::make_shared<copra::TrajectoryCost>(Eigen::MatrixXd::Identity(2, 2), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addCost(badP), std::domain_error);
        auto badT = std::make_shared<copra::TargetCost>(Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(2));
        REQUIRE_THROWS_AS(controller.addCost(badT), std::domain_error);
        auto badT2 = std::make_shared<copra::TargetCost>(Eigen::MatrixXd::Identity(2, 2), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addCost(badT2), std::domain_error);
        auto badC = std::make_shared<copra::ControlCost>(Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addCost(badC), std::domain_error);
        auto badC2 = std::make_shared<copra::ControlCost>(Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addCost(badC2), std::domain_error);
        auto badX1 = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Identity(5, 5), Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addCost(badX1), std::domain_error);
        auto badX2 = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addCost(badX2), std::domain_error);
        auto badX3 = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addCost(badX3), std::domain_error);
    }
    { // constraints
        auto t1 = std::make_shared<copra::TrajectoryConstraint>(Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(2));
        REQUIRE_THROWS_AS(controller.addConstraint(t1), std::domain_error);
        auto t2 = std::make_shared<copra::TrajectoryConstraint>(Eigen::MatrixXd::Identity(2, 2), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addConstraint(t2), std::domain_error);
        auto c1 = std::make_shared<copra::ControlConstraint>(Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addConstraint(c1), std::domain_error);
        auto c2 = std::make_shared<copra::ControlConstraint>(Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addConstraint(c2), std::domain_error);
        auto good = std::make_shared<copra::ControlConstraint>(s.G, s.h);
        REQUIRE_NOTHROW(controller.addConstraint(good));
        REQUIRE_THROWS_AS(controller.addConstraint(good), std::runtime_error); // move semantics: single initialisation
        auto m1 = std::make_shared<copra::MixedConstraint>(Eigen::MatrixXd::Identity(5, 5), Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addConstraint(m1), std::domain_error);
        auto m2 = std::make_shared<copra::MixedConstraint>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Identity(5, 5), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addConstraint(m2), std::domain_error);
        auto m3 = std::make_shared<copra::MixedConstraint>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Identity(1, 1), Eigen::VectorXd::Ones(5));
        REQUIRE_THROWS_AS(controller.addConstraint(m3), std::domain_error);
    }
    { // bounds
        BoundedSystem b;
        auto tb1 = std::make_shared<copra::TrajectoryBoundConstraint>(Eigen::VectorXd::Ones(3), Eigen::VectorXd::Ones(2));
        REQUIRE_THROWS_AS(controller.addConstraint(tb1), std::domain_error);
        auto tb2 = std::make_shared<copra::TrajectoryBoundConstraint>(Eigen::VectorXd::Ones(2), Eigen::VectorXd::Ones(3));
        REQUIRE_THROWS_AS(controller.addConstraint(tb2), std::domain_error);
        auto tb3 = std::make_shared<copra::TrajectoryBoundConstraint>(Eigen::VectorXd::Ones(3), Eigen::VectorXd::Ones(3));
        REQUIRE_THROWS_AS(controller.addConstraint(tb3), std::domain_error);
        auto cb1 = std::make_shared<copra::ControlBoundConstraint>(Eigen::VectorXd::Ones(3), Eigen::VectorXd::Ones(1));
        REQUIRE_THROWS_AS(controller.addConstraint(cb1), std::domain_error);
        auto cb2 = std::make_shared<copra::ControlBoundConstraint>(Eigen::VectorXd::Ones(3), Eigen::VectorXd::Ones(3));
        REQUIRE_THROWS_AS(controller.addConstraint(cb2), std::domain_error);
        auto good = std::make_shared<copra::ControlBoundConstraint>(b.uLower, b.uUpper);
        REQUIRE_NOTHROW(controller.addConstraint(good));
        REQUIRE_THROWS_AS(controller.addConstraint(good), std::runtime_error);
    }
}

static void autoSpanShapes() // tests/TestLMPC.cpp:777-943: every mix of step-size and pre-spanned inputs is accepted
{
    {
        BoundedSystem s;
        auto ps = std::make_shared<copra::PreviewSystem>();
        ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
        auto controller = copra::LMPC(ps);
        const int nx = s.nbStep + 1;
        auto check = [&](const Eigen::VectorXd& xl, const Eigen::VectorXd& xu, const Eigen::VectorXd& ul, const Eigen::VectorXd& uu) {
            auto tc = std::make_shared<copra::TrajectoryBoundConstraint>(xl, xu);
            tc->autoSpan();
            auto cc = std::make_shared<copra::ControlBoundConstraint>(ul, uu);
            cc->autoSpan();
            REQUIRE_NOTHROW(controller.addConstraint(tc));
            REQUIRE_NOTHROW(controller.addConstraint(cc));
        };
        check(s.xLower, s.xUpper, s.uLower, s.uUpper);
        check(spanVec(s.xLower, nx), s.xUpper, spanVec(s.uLower, s.nbStep), s.uUpper);
        check(s.xLower, spanVec(s.xUpper, nx), s.uLower, spanVec(s.uUpper, s.nbStep));
        check(spanVec(s.xLower, nx), spanVec(s.xUpper, nx), spanVec(s.uLower, s.nbStep), spanVec(s.uUpper, s.nbStep));
    }
    {
        IneqSystem s;
        auto ps = std::make_shared<copra::PreviewSystem>();
        ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
        auto controller = copra::LMPC(ps);
        const int nx = s.nbStep + 1;
        auto check = [&](const Eigen::MatrixXd& E, const Eigen::VectorXd& p, const Eigen::MatrixXd& G, const Eigen::VectorXd& h) {
            auto tc = std::make_shared<copra::TrajectoryConstraint>(E, p);
            tc->autoSpan();
            auto cc = std::make_shared<copra::ControlConstraint>(G, h);
            cc->autoSpan();
            REQUIRE_NOTHROW(controller.addConstraint(tc));
            REQUIRE_NOTHROW(controller.addConstraint(cc));
        };
        check(s.E, s.p, s.G, s.h);
        check(spanMat(s.E, nx), s.p, spanMat(s.G, s.nbStep), s.h);
        check(s.E, spanVec(s.p, nx), s.G, spanVec(s.h, s.nbStep));
        check(spanMat(s.E, nx), spanVec(s.p, nx), spanMat(s.G, s.nbStep), spanVec(s.h, s.nbStep));
        auto checkCost = [&](const Eigen::MatrixXd& M, const Eigen::VectorXd& p, const Eigen::VectorXd& w) {
            auto cost = std::make_shared<copra::TrajectoryCost>(M, p);
            cost->weights(w);
            cost->autoSpan();
            REQUIRE_NOTHROW(controller.addCost(cost));
        };
        checkCost(s.M, s.xd, s.wx);
        checkCost(spanMat(s.M, nx), s.xd, s.wx);
        checkCost(s.M, spanVec(s.xd, nx), s.wx);
        checkCost(spanMat(s.M, nx), spanVec(s.xd, nx), s.wx);
        auto checkCtrl = [&](const Eigen::MatrixXd& N, const Eigen::VectorXd& p, const Eigen::VectorXd& w) {
            auto cost = std::make_shared<copra::ControlCost>(N, p);
            cost->weights(w);
            cost->autoSpan();
            REQUIRE_NOTHROW(controller.addCost(cost));
        };
        checkCtrl(s.N, s.ud, s.wu);
        checkCtrl(spanMat(s.N, s.nbStep), s.ud, s.wu);
        checkCtrl(s.N, spanVec(s.ud, s.nbStep), s.wu);
        auto checkMixed = [&](const Eigen::MatrixXd& M, const Eigen::MatrixXd& N, const Eigen::VectorXd& p) {
            auto cost = std::make_shared<copra::MixedCost>(M, N, p);
            cost->autoSpan();
            REQUIRE_NOTHROW(controller.addCost(cost));
        };
        const Eigen::MatrixXd Mm = Eigen::MatrixXd::Ones(1, 2), Nm = Eigen::MatrixXd::Ones(1, 1);
        const Eigen::VectorXd pm = Eigen::VectorXd::Ones(1);
        checkMixed(Mm, Nm, pm);
        checkMixed(spanMat(Mm, s.nbStep, 1), Nm, pm);
        checkMixed(Mm, spanMat(Nm, s.nbStep), pm);
        checkMixed(Mm, Nm, spanVec(pm, s.nbStep));
    }
    {
        MixedSystem s;
        auto ps = std::make_shared<copra::PreviewSystem>();
        ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
        auto controller = copra::LMPC(ps);
        auto check = [&](const Eigen::MatrixXd& E, const Eigen::MatrixXd& G, const Eigen::VectorXd& p) {
            auto mc = std::make_shared<copra::MixedConstraint>(E, G, p);
            mc->autoSpan();
            REQUIRE_NOTHROW(controller.addConstraint(mc));
        };
        check(s.E, s.G, s.p);
        check(spanMat(s.E, s.nbStep, 1), s.G, s.p);
        check(s.E, spanMat(s.G, s.nbStep), s.p);
        check(s.E, s.G, spanVec(s.p, s.nbStep));
    }
    { // AutoSpan itself
        Eigen::MatrixXd m = Eigen::MatrixXd::Ones(2, 3);
        copra::AutoSpan::spanMatrix(m, 6, 1);
        REQUIRE(m.rows() == 6 && m.cols() == 12 && m(2, 3) == 1.0 && m(2, 0) == 0.0);
        Eigen::VectorXd v = Eigen::VectorXd::Ones(2);
        REQUIRE_THROWS_AS(copra::AutoSpan::spanVector(v, 5), std::domain_error);
    }
}

// ------------------------------------------------------------------------------------------------
// GPU groups
// ------------------------------------------------------------------------------------------------
static void quadProgProblem() // tests/TestSolvers.cpp:25-33 through the new backend, plus the known answer KA-1
{
    QpProblem q;
    copra::B200Solver solver;
    solver.SI_problem(q.nrvars, q.nreqs, q.nrineqs);
    REQUIRE(solver.SI_solve(q.Q, q.c, q.Aeq, q.beq, q.Aineq, q.bineq, q.XL, q.XU));
    REQUIRE(solver.SI_fail() == 0);
    const double xs[6] = { 1.797542603546133, -0.338148723827743, 0.163388028104626, -4.988402270313609, 0.605494327732570, -3.115562338676213 };
    for (int i = 0; i < 6; ++i) REQUIRE(std::fabs(solver.SI_result()(i) - xs[i]) < 1e-12);
    REQUIRE(solver.SI_iter() == 5);
    auto viaFactory = copra::solverFactory(copra::SolverFlag::QuadProgDense);
    viaFactory->SI_problem(q.nrvars, q.nreqs, q.nrineqs);
    REQUIRE(viaFactory->SI_solve(q.Q, q.c, q.Aeq, q.beq, q.Aineq, q.bineq, q.XL, q.XU));
}

template <class CostT> static void boundedCase(bool mixedCost) // tests/TestLMPC.cpp:36-213
{
    BoundedSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    std::shared_ptr<copra::CostFunction> xCost, uCost;
    if (mixedCost) {
        xCost = std::make_shared<copra::MixedCost>(s.M, Eigen::MatrixXd::Zero(2, 1), s.xd);
        uCost = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Zero(1, 2), s.N, s.ud);
    } else {
        xCost = std::make_shared<CostT>(s.M, s.xd);
        uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
    }
    auto trajConstr = std::make_shared<copra::TrajectoryBoundConstraint>(s.xLower, s.xUpper);
    auto contConstr = std::make_shared<copra::ControlBoundConstraint>(s.uLower, s.uUpper);
    xCost->weights(s.wx);
    uCost->weights(s.wu);
    controller.addCost(xCost);
    controller.addCost(uCost);
    controller.addConstraint(trajConstr);
    controller.addConstraint(contConstr);
    REQUIRE(controller.solve());
    Eigen::VectorXd pos, vel;
    splitTraj(controller.trajectory(), pos, vel);
    const Eigen::VectorXd control = controller.control();
    REQUIRE_LE(std::fabs(s.xd(1) - vel.tail(mixedCost ? 3 : 1)(0)), 0.001);
    REQUIRE_LE(pos.maxCoeff(), s.x0(0));
    REQUIRE_LE(vel.maxCoeff(), s.xUpper(1) + 1e-6);
    REQUIRE_LE(control.maxCoeff(), s.uUpper(0) + 1e-6);
    REQUIRE(controller.solveAndBuildTime() >= controller.solveTime());
    REQUIRE(controller.nrIneqConstr() == s.nbStep + 1 && controller.nrEqConstr() == 0);
    REQUIRE(controller.Q().rows() == s.nbStep && controller.Aineq().rows() == s.nbStep + 1 && controller.ub()(0) == 200.0);
}

template <class CostT> static void ineqCase() // tests/TestLMPC.cpp:219-409
{
    IneqSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    auto xCost = std::make_shared<CostT>(s.M, s.xd);
    auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
    auto trajConstr = std::make_shared<copra::TrajectoryConstraint>(s.E, s.p);
    auto contConstr = std::make_shared<copra::ControlConstraint>(s.G, s.h);
    xCost->weights(s.wx);
    uCost->weights(s.wu);
    controller.addCost(xCost);
    controller.addCost(uCost);
    controller.addConstraint(trajConstr);
    controller.addConstraint(contConstr);
    REQUIRE(controller.solve());
    Eigen::VectorXd pos, vel;
    splitTraj(controller.trajectory(), pos, vel);
    REQUIRE_LE(std::fabs(s.xd(1) - vel.tail(1)(0)), 0.001);
    REQUIRE_LE(pos.maxCoeff(), s.x0(0));
    REQUIRE_LE(vel.maxCoeff(), s.p(0) + 1e-6);
    REQUIRE_LE(controller.control().maxCoeff(), s.h(0) + 1e-6);
}

template <class CostT> static void mixedCase() // tests/TestLMPC.cpp:415-587
{
    MixedSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    auto xCost = std::make_shared<CostT>(s.M, s.xd);
    auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
    auto mixedConstr = std::make_shared<copra::MixedConstraint>(s.E, s.G, s.p);
    xCost->weights(s.wx);
    uCost->weights(s.wu);
    controller.addCost(xCost);
    controller.addCost(uCost);
    controller.addConstraint(mixedConstr);
    REQUIRE(controller.solve());
    Eigen::VectorXd pos, vel;
    const Eigen::VectorXd traj = controller.trajectory(), control = controller.control();
    splitTraj(traj, pos, vel);
    REQUIRE_LE(std::fabs(s.xd(1) - vel.tail(1)(0)), 0.001);
    REQUIRE_LE(pos.maxCoeff(), s.x0(0));
    bool ok = true;
    for (int i = 0; i < s.nbStep; ++i) {
        const double res = s.E(0, 0) * traj(2 * i) + s.E(0, 1) * traj(2 * i + 1) + s.G(0, 0) * control(i);
        if (!(res <= s.p(0) + 1e-6)) ok = false;
    }
    REQUIRE(ok);
}

template <class CostT> static void eqCase() // tests/TestLMPC.cpp:593-771
{
    EqSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    auto xCost = std::make_shared<CostT>(s.M, s.xd);
    auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
    auto trajConstr = std::make_shared<copra::TrajectoryConstraint>(s.E, s.p, false);
    xCost->weights(s.wx);
    uCost->weights(s.wu);
    controller.addCost(xCost);
    controller.addCost(uCost);
    controller.addConstraint(trajConstr);
    REQUIRE(controller.solve());
    Eigen::VectorXd pos, vel;
    splitTraj(controller.trajectory(), pos, vel);
    REQUIRE_LE(pos.maxCoeff(), s.x0(0) + 1e-6);
    REQUIRE_LE(vel.maxCoeff(), s.p(0) + 1e-6);
    REQUIRE(controller.nrEqConstr() == 2 * (s.nbStep + 1));
}

static void removeAndDelete() // tests/TestLMPC.cpp:1093-1118 + the use_count semantics of src/LMPC.cpp:288-307
{
    IneqSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>(s.A, s.B, s.c, s.x0, s.nbStep);
    auto controller = copra::LMPC(ps);
    {
        auto xCost = std::make_shared<copra::TargetCost>(s.M, s.xd);
        auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
        auto trajConstr = std::make_shared<copra::TrajectoryConstraint>(s.E, s.p);
        auto contConstr = std::make_shared<copra::ControlConstraint>(s.G, s.h);
        controller.addCost(xCost);
        controller.addCost(uCost);
        controller.addConstraint(trajConstr);
        controller.addConstraint(contConstr);
        controller.removeCost(xCost);
        controller.removeCost(uCost);
        controller.removeConstraint(trajConstr);
        controller.removeConstraint(contConstr);
    }
    REQUIRE(controller.solve()); // only the 1e-6 I regulariser is left: U = 0
    REQUIRE(controller.nrIneqConstr() == 0);
    REQUIRE(std::fabs(controller.control().maxCoeff()) < 1e-12);
    {
        auto dropped = std::make_shared<copra::TrajectoryConstraint>(s.E, s.p);
        controller.addConstraint(dropped);
    } // the user's shared_ptr is gone: removed after the next solve
    REQUIRE(controller.solve());
    REQUIRE(controller.nrIneqConstr() == s.nbStep + 1);
    REQUIRE(controller.solve());
    REQUIRE(controller.nrIneqConstr() == 0);
}

// a user-supplied SolverInterface (the reference's plug-in protocol): here it simply forwards to B200Solver
class ForwardingSolver : public copra::SolverInterface {
public:
    int SI_fail() const override { return inner_.SI_fail(); }
    void SI_inform() const override { inner_.SI_inform(); }
    const Eigen::VectorXd& SI_result() const override { return inner_.SI_result(); }
    void SI_problem(int a, int b, int c) override { ++problems; inner_.SI_problem(a, b, c); }
    bool SI_solve(const Eigen::MatrixXd& Q, const Eigen::VectorXd& c, const Eigen::MatrixXd& Aeq, const Eigen::VectorXd& beq, const Eigen::MatrixXd& Aineq,
        const Eigen::VectorXd& bineq, const Eigen::VectorXd& XL, const Eigen::VectorXd& XU) override
    {
        ++solves;
        return inner_.SI_solve(Q, c, Aeq, beq, Aineq, bineq, XL, XU);
    }
    int problems = 0, solves = 0;

private:
    copra::B200Solver inner_;
};

static void pluginProtocolAndGetters()
{
    BoundedSystem s;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto fused = copra::LMPC(ps);
    auto plugged = copra::LMPC(ps);
    auto fwd = new ForwardingSolver();
    plugged.useSolver(std::unique_ptr<copra::SolverInterface>(fwd));
    auto xCost = std::make_shared<copra::TargetCost>(s.M, s.xd);
    auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
    auto trajConstr = std::make_shared<copra::TrajectoryBoundConstraint>(s.xLower, s.xUpper);
    auto contConstr = std::make_shared<copra::ControlBoundConstraint>(s.uLower, s.uUpper);
    auto contConstr2 = std::make_shared<copra::ControlBoundConstraint>(s.uLower, s.uUpper);
    xCost->weights(s.wx);
    uCost->weights(s.wu);
    for (auto* ctl : { &fused, &plugged }) {
        ctl->addCost(xCost);
        ctl->addCost(uCost);
        ctl->addConstraint(trajConstr);
    }
    fused.addConstraint(contConstr);
    plugged.addConstraint(contConstr2);
    REQUIRE(fused.solve());
    REQUIRE(plugged.solve());
    REQUIRE(fwd->problems == 1 && fwd->solves == 1);
    REQUIRE(fused.control().isApprox(plugged.control(), 1e-9));
    REQUIRE(fused.trajectory().isApprox(plugged.trajectory(), 1e-9));
    // per-object getters (CostFunction::Q/c/E/f, EqIneqConstraint::A/b/Y/z) evaluated on the GPU
    xCost->update(*ps);
    uCost->update(*ps);
    trajConstr->update(*ps);
    const Eigen::MatrixXd& Q = fused.Q();
    bool same = true;
    for (int j = 0; j < s.nbStep; ++j)
        for (int i = 0; i < s.nbStep; ++i) {
            const double want = (i == j ? 1e-6 : 0.0) + xCost->Q()(i, j) + uCost->Q()(i, j);
            if (std::fabs(Q(i, j) - want) > 1e-12 * std::max(1.0, std::fabs(want))) same = false;
        }
    REQUIRE(same);
    REQUIRE(trajConstr->A().rows() == s.nbStep + 1 && trajConstr->Y().cols() == 2);
    bool bsame = true;
    for (int i = 0; i < s.nbStep + 1; ++i) {
        const double bi = trajConstr->z()(i) - (trajConstr->Y()(i, 0) * s.x0(0) + trajConstr->Y()(i, 1) * s.x0(1));
        if (std::fabs(bi - trajConstr->b()(i)) > 1e-12) bsame = false;
        if (std::fabs(trajConstr->b()(i) - fused.bineq()(i)) > 0.0) bsame = false;
    }
    REQUIRE(bsame);
}

static void initialStateComparison() // tests/TestLMPC_InitialState.cpp:29-260 (step-size entries, QuadProg-equivalent backend)
{
    const int xDim = 2, uDim = 1, N = 10;
    Eigen::MatrixXd A = Eigen::MatrixXd::Ones(xDim, xDim), B = Eigen::MatrixXd::Ones(xDim, uDim);
    Eigen::VectorXd d = Eigen::VectorXd::Zero(xDim), x0 = Eigen::VectorXd::Zero(xDim);
    A(1, 0) = 0.0; A(0, 1) = 0.1; B(0, 0) = 0.005;
    x0(1) = -1.0;
    auto ps = std::make_shared<copra::PreviewSystem>(A, B, d, x0, N);
    auto mk = [&](copra::LMPC& ctl) {
        auto tc = std::make_shared<copra::TrajectoryCost>(Eigen::MatrixXd::Identity(xDim, xDim), Eigen::VectorXd::Zero(xDim));
        auto gc = std::make_shared<copra::TargetCost>(Eigen::MatrixXd::Identity(xDim, xDim), Eigen::VectorXd::Ones(xDim));
        auto cc = std::make_shared<copra::ControlCost>(Eigen::MatrixXd::Identity(uDim, uDim), Eigen::VectorXd::Zero(uDim));
        auto mc = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Ones(1, xDim), Eigen::MatrixXd::Ones(1, uDim), Eigen::VectorXd::Zero(1));
        tc->weight(0.5); gc->weight(2.0); cc->weight(0.1); mc->weight(0.3);
        auto tcs = std::make_shared<copra::TrajectoryConstraint>(Eigen::MatrixXd::Identity(xDim, xDim), Eigen::VectorXd::Constant(xDim, 50.0));
        auto ccs = std::make_shared<copra::ControlConstraint>(Eigen::MatrixXd::Identity(uDim, uDim), Eigen::VectorXd::Constant(uDim, 20.0));
        auto mcs = std::make_shared<copra::MixedConstraint>(Eigen::MatrixXd::Ones(1, xDim), Eigen::MatrixXd::Ones(1, uDim), Eigen::VectorXd::Constant(1, 60.0));
        auto cbs = std::make_shared<copra::ControlBoundConstraint>(Eigen::VectorXd::Constant(uDim, -10.0), Eigen::VectorXd::Constant(uDim, 10.0));
        std::vector<std::shared_ptr<copra::CostFunction>> costs = { tc, gc, cc, mc };
        std::vector<std::shared_ptr<copra::Constraint>> cstrs = { tcs, ccs, mcs, cbs };
        for (auto& c : costs) ctl.addCost(c);
        for (auto& c : cstrs) ctl.addConstraint(c);
        return std::make_pair(costs, cstrs);
    };
    copra::LMPC lmpc(ps);
    copra::InitialStateLMPC islmpc(ps); // default bounds pin x0 (quirk Q7)
    // The reference runs this comparison with QLD and the default R = 0, which makes the [x0; U] Hessian
    // singular (its Schur complement is R): a Goldfarb-Idnani / QuadProg backend reports fail = 2 there.
    // With x0 pinned R does not change the solution, so a positive definite R keeps the comparison exact.
    islmpc.resetInitialStateCost(Eigen::MatrixXd::Identity(xDim, xDim), Eigen::VectorXd::Zero(xDim));
    auto keep1 = mk(lmpc);
    auto keep2 = mk(islmpc);
    REQUIRE(lmpc.solve());
    REQUIRE(islmpc.solve());
    const int n = uDim * N;
    const Eigen::MatrixXd &Q1 = lmpc.Q(), &Q2 = islmpc.Q(), &A1 = lmpc.Aineq(), &A2 = islmpc.Aineq();
    REQUIRE(Q2.rows() == xDim + n && A2.cols() == xDim + n && A1.rows() == A2.rows());
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) if (std::fabs(Q1(i, j) - Q2(xDim + i, xDim + j)) > 1e-6) ok = false;
        for (int i = 0; i < A1.rows(); ++i) if (std::fabs(A1(i, j) - A2(i, xDim + j)) > 1e-6) ok = false;
        if (std::fabs(lmpc.lb()(j) - islmpc.lb()(xDim + j)) > 1e-6 || std::fabs(lmpc.ub()(j) - islmpc.ub()(xDim + j)) > 1e-6) ok = false;
    }
    REQUIRE(ok);
    const Eigen::VectorXd xi0 = islmpc.initialState();
    REQUIRE(std::fabs(xi0(0) - x0(0)) < 1e-9 && std::fabs(xi0(1) - x0(1)) < 1e-9);
    REQUIRE(lmpc.control().isApprox(islmpc.control(), 1e-6));
    // optimisation variant (:266-396): x0 box +-1, R = 1e-6 I
    islmpc.resetInitialStateBounds(Eigen::VectorXd::Constant(xDim, -1.0), Eigen::VectorXd::Constant(xDim, 1.0));
    Eigen::MatrixXd R = Eigen::MatrixXd::Identity(xDim, xDim);
    R *= 1e-6;
    islmpc.resetInitialStateCost(R, Eigen::VectorXd::Zero(xDim));
    REQUIRE(islmpc.solve());
    const Eigen::VectorXd xo = islmpc.initialState();
    for (int i = 0; i < xDim; ++i) { REQUIRE_LE(xo(i), 1.0 + 1e-6); REQUIRE_LE(-1.0 - 1e-6, xo(i)); }
}

static void fullSizeEntriesSolve() // autoSpan'd (full-size) entries give the same controller as step-size entries
{
    IneqSystem s;
    s.nbStep = 40;
    auto ps = std::make_shared<copra::PreviewSystem>();
    ps->system(s.A, s.B, s.c, s.x0, s.nbStep);
    auto build = [&](copra::LMPC& ctl, bool span) {
        auto xCost = std::make_shared<copra::TrajectoryCost>(s.M, s.xd);
        auto uCost = std::make_shared<copra::ControlCost>(s.N, s.ud);
        auto mCost = std::make_shared<copra::MixedCost>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Ones(1, 1), Eigen::VectorXd::Ones(1));
        auto trajConstr = std::make_shared<copra::TrajectoryConstraint>(s.E, s.p);
        auto contConstr = std::make_shared<copra::ControlConstraint>(s.G, s.h);
        auto mixConstr = std::make_shared<copra::MixedConstraint>(Eigen::MatrixXd::Ones(1, 2), Eigen::MatrixXd::Ones(1, 1), Eigen::VectorXd::Constant(1, 500.0));
        auto bound = std::make_shared<copra::ControlBoundConstraint>(Eigen::VectorXd::Constant(1, -300.0), Eigen::VectorXd::Constant(1, 250.0));
        xCost->weights(s.wx);
        uCost->weights(s.wu);
        mCost->weight(1e-3);
        if (span) { xCost->autoSpan(); uCost->autoSpan(); mCost->autoSpan(); trajConstr->autoSpan(); contConstr->autoSpan(); mixConstr->autoSpan(); bound->autoSpan(); }
        ctl.addCost(xCost); ctl.addCost(uCost); ctl.addCost(mCost);
        ctl.addConstraint(trajConstr); ctl.addConstraint(contConstr); ctl.addConstraint(mixConstr); ctl.addConstraint(bound);
        std::vector<std::shared_ptr<void>> keep = { xCost, uCost, mCost, trajConstr, contConstr, mixConstr, bound };
        return keep;
    };
    copra::LMPC stepSize(ps), fullSize(ps);
    auto k1 = build(stepSize, false);
    auto k2 = build(fullSize, true);
    // autoSpan() of step-size inputs leaves them step-size (max_dim == rows): span by hand to get real full-size entries
    auto fM = std::make_shared<copra::TrajectoryCost>(spanMat(s.M, s.nbStep + 1), spanVec(s.xd, s.nbStep + 1));
    fM->weights(s.wx);
    auto fE = std::make_shared<copra::TrajectoryConstraint>(spanMat(s.E, s.nbStep + 1), spanVec(s.p, s.nbStep + 1));
    auto fG = std::make_shared<copra::ControlConstraint>(spanMat(s.G, s.nbStep), spanVec(s.h, s.nbStep));
    copra::LMPC dense(ps);
    auto uCost = std::make_shared<copra::ControlCost>(spanMat(s.N, s.nbStep), spanVec(s.ud, s.nbStep));
    uCost->weights(s.wu);
    auto mCost = std::make_shared<copra::MixedCost>(spanMat(Eigen::MatrixXd::Ones(1, 2), s.nbStep, 1), spanMat(Eigen::MatrixXd::Ones(1, 1), s.nbStep), Eigen::VectorXd::Ones(s.nbStep));
    mCost->weight(1e-3);
    auto mixConstr = std::make_shared<copra::MixedConstraint>(spanMat(Eigen::MatrixXd::Ones(1, 2), s.nbStep, 1), spanMat(Eigen::MatrixXd::Ones(1, 1), s.nbStep), Eigen::VectorXd::Constant(s.nbStep, 500.0));
    auto bound = std::make_shared<copra::ControlBoundConstraint>(Eigen::VectorXd::Constant(s.nbStep, -300.0), Eigen::VectorXd::Constant(s.nbStep, 250.0));
    dense.addCost(fM); dense.addCost(uCost); dense.addCost(mCost);
    dense.addConstraint(fE); dense.addConstraint(fG); dense.addConstraint(mixConstr); dense.addConstraint(bound);
    REQUIRE(fM->fullSizeEntry() && fE->fullSizeEntry() && fG->fullSizeEntry() && mCost->fullSizeEntry() && bound->fullSizeEntry());
    REQUIRE(stepSize.solve());
    REQUIRE(fullSize.solve());
    REQUIRE(dense.solve());
    REQUIRE(stepSize.control().isApprox(fullSize.control(), 1e-9));
    REQUIRE(stepSize.control().isApprox(dense.control(), 1e-7));
    REQUIRE(stepSize.trajectory().isApprox(dense.trajectory(), 1e-7));
    REQUIRE(stepSize.nrIneqConstr() == dense.nrIneqConstr());
    const Eigen::MatrixXd &Qa = stepSize.Q(), &Qb = dense.Q();
    bool same = true;
    for (int j = 0; j < s.nbStep; ++j)
        for (int i = 0; i < s.nbStep; ++i)
            if (std::fabs(Qa(i, j) - Qb(i, j)) > 1e-10 * std::max(1.0, std::fabs(Qa(i, j)))) same = false;
    REQUIRE(same);
}

static void batchedEntry()
{
    BoundedSystem s;
    const int batch = 64, N = 50;
    std::vector<double> x0(2 * batch), xd(2 * batch), uup(batch);
    for (int b = 0; b < batch; ++b) {
        x0[2 * b] = 0.0; x0[2 * b + 1] = -5.0 + 0.01 * b;
        xd[2 * b] = 0.0; xd[2 * b + 1] = -1.0 - 0.005 * b;
        uup[b] = 180.0 + b;
    }
    Eigen::MatrixXd A(2, 2), B(2, 1);
    Eigen::VectorXd d(2);
    const double T = 0.03;
    A << 1, T, 0, 1;
    B << 0.5 * T * T / s.mass, T / s.mass;
    d << (-9.81 / 2.) * T * T, -9.81 * T;
    copra::BatchedLMPC ctl(2, 1, N, batch);
    ctl.system(copra::b200::arr(A.data()), copra::b200::arr(B.data()), copra::b200::arr(d.data()), copra::b200::arr(x0.data(), 2));
    copra_b200_cost target{};
    target.kind = COPRA_B200_COST_TARGET; target.rows = 2;
    target.M = copra::b200::arr(s.M.data()); target.p = copra::b200::arr(xd.data(), 2); target.w = copra::b200::arr(s.wx.data());
    copra_b200_cost effort{};
    effort.kind = COPRA_B200_COST_CONTROL; effort.rows = 1;
    effort.N = copra::b200::arr(s.N.data()); effort.p = copra::b200::arr(s.ud.data()); effort.w = copra::b200::arr(s.wu.data());
    copra_b200_constraint tb{};
    tb.kind = COPRA_B200_CSTR_TRAJECTORY_BOUND; tb.rows = 2;
    tb.lower = copra::b200::arr(s.xLower.data()); tb.upper = copra::b200::arr(s.xUpper.data());
    copra_b200_constraint cb{};
    cb.kind = COPRA_B200_CSTR_CONTROL_BOUND; cb.rows = 1;
    cb.lower = copra::b200::arr(s.uLower.data()); cb.upper = copra::b200::arr(uup.data(), 1);
    ctl.addCost(target);
    ctl.addCost(effort);
    ctl.addConstraint(tb);
    ctl.addConstraint(cb);
    REQUIRE(ctl.solve() == batch);
    // instance 7 against a single-instance controller
    Eigen::VectorXd x07(2), xd7(2), uu7(1);
    x07 << x0[14], x0[15];
    xd7 << xd[14], xd[15];
    uu7 << uup[7];
    auto ps = std::make_shared<copra::PreviewSystem>(A, B, d, x07, N);
    copra::LMPC one(ps);
    auto c1 = std::make_shared<copra::TargetCost>(s.M, xd7);
    auto c2 = std::make_shared<copra::ControlCost>(s.N, s.ud);
    c1->weights(s.wx); c2->weights(s.wu);
    auto k1 = std::make_shared<copra::TrajectoryBoundConstraint>(s.xLower, s.xUpper);
    auto k2 = std::make_shared<copra::ControlBoundConstraint>(s.uLower, uu7);
    one.addCost(c1); one.addCost(c2); one.addConstraint(k1); one.addConstraint(k2);
    REQUIRE(one.solve());
    bool same = true;
    for (int i = 0; i < N; ++i) if (ctl.controls()[size_t(7) * N + i] != one.control()(i)) same = false;
    REQUIRE(same); // instances are independent: bit-identical to the batch-of-one result
    REQUIRE(ctl.solveAndBuildTime() > 0 && ctl.solveTime() > 0);
    // receding-horizon step: new initial states on the resident build == a fresh controller with those states
    std::vector<double> x1(x0);
    for (int b = 0; b < batch; ++b) x1[2 * b + 1] += 0.25;
    REQUIRE(ctl.solve() == batch);                                  // the engine holds this batch again ...
    REQUIRE(ctl.resolve(copra::b200::arr(x1.data(), 2)) == batch); // ... so this is the K4 + K5..K7 path
    REQUIRE(ctl.solveTime() > 0);
    x07 << x1[14], x1[15];
    ps->xInit(x07);
    REQUIRE(one.solve());
    same = true;
    for (int i = 0; i < N; ++i) if (ctl.controls()[size_t(7) * N + i] != one.control()(i)) same = false;
    REQUIRE(same);
    // data-parallel sharding: the same batch over three engine handles (device 0 thrice when the box has one
    // GPU; distinct devices otherwise) -- ragged shards of 22 / 22 / 20 instances, results BITWISE the single-handle ones
    const std::vector<double> ref_u = ctl.controls(), ref_x = ctl.trajectories();
    const std::vector<int> ref_it = ctl.iterations();
    const int ndev = copra_b200_device_count();
    ctl.useDevices(ndev >= 3 ? std::vector<int>{ 0, 1, 2 } : std::vector<int>{ 0, 0, 0 });
    REQUIRE(ctl.nrDevices() == 3);
    REQUIRE(ctl.solve() == batch);                                  // full solve of the x0 batch, sharded
    REQUIRE(ctl.resolve(copra::b200::arr(x1.data(), 2)) == batch); // sharded receding-horizon step
    REQUIRE(ctl.controls() == ref_u);
    REQUIRE(ctl.trajectories() == ref_x);
    REQUIRE(ctl.iterations() == ref_it);
}

// Two host threads, each with its own controllers: every thread gets its own engine handle (b200::handle() is thread-local),
// so the solves run concurrently and a controller solved on one thread serves its getters from a snapshot / rebuild on another.
static void twoHostThreads()
{
    BoundedSystem s;
    auto solveOne = [&s](double ud, Eigen::VectorXd* out, Eigen::MatrixXd* Qout, bool* ok) {
        auto ps = std::make_shared<copra::PreviewSystem>();
        ps->system(s.A, s.B, s.c, s.x0, 60);
        copra::LMPC ctl(ps);
        auto xCost = std::make_shared<copra::TargetCost>(s.M, s.xd);
        Eigen::VectorXd udv(1);
        udv << ud;
        auto uCost = std::make_shared<copra::ControlCost>(s.N, udv);
        auto trajConstr = std::make_shared<copra::TrajectoryBoundConstraint>(s.xLower, s.xUpper);
        auto contConstr = std::make_shared<copra::ControlBoundConstraint>(s.uLower, s.uUpper);
        xCost->weights(s.wx);
        uCost->weights(s.wu);
        ctl.addCost(xCost); ctl.addCost(uCost); ctl.addConstraint(trajConstr); ctl.addConstraint(contConstr);
        *ok = true;
        for (int rep = 0; rep < 5; ++rep) *ok = *ok && ctl.solve();
        *out = ctl.control();
        *Qout = ctl.Q();
    };
    Eigen::VectorXd u1, u2, r1, r2;
    Eigen::MatrixXd Q1, Q2, Qr1, Qr2;
    bool ok1 = false, ok2 = false, okr1 = false, okr2 = false;
    solveOne(2.0, &r1, &Qr1, &okr1); // single-threaded references
    solveOne(3.0, &r2, &Qr2, &okr2);
    std::thread t1([&] { solveOne(2.0, &u1, &Q1, &ok1); });
    std::thread t2([&] { solveOne(3.0, &u2, &Q2, &ok2); });
    t1.join();
    t2.join();
    REQUIRE(ok1 && ok2 && okr1 && okr2);
    auto same = [](const double* a, const double* b, long n) { for (long i = 0; i < n; ++i) if (a[i] != b[i]) return false; return true; };
    REQUIRE(u1.size() == r1.size() && same(u1.data(), r1.data(), long(r1.size())));
    REQUIRE(u2.size() == r2.size() && same(u2.data(), r2.data(), long(r2.size())));
    REQUIRE(Q1.size() == Qr1.size() && same(Q1.data(), Qr1.data(), long(Qr1.size())));
    REQUIRE(Q2.size() == Qr2.size() && same(Q2.data(), Qr2.data(), long(Qr2.size())));
    REQUIRE(!same(u1.data(), u2.data(), long(u1.size())));
}

int main(int argc, char** argv)
{
    std::setvbuf(stdout, nullptr, _IONBF, 0);
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    std::vector<std::pair<std::string, std::function<void()>>> groups = {
        { "ERROR_HANDLERS", errorHandlers }, { "AUTOSPAN_SHAPES", autoSpanShapes } };
    if (gpu) {
        groups.push_back({ "QUADPROG_PROBLEM_B200", quadProgProblem });
        groups.push_back({ "MPC_TARGET_COST_WITH_BOUND_CONSTRAINTS", [] { boundedCase<copra::TargetCost>(false); } });
        groups.push_back({ "MPC_TRAJECTORY_COST_WITH_BOUND_CONSTRAINTS", [] { boundedCase<copra::TrajectoryCost>(false); } });
        groups.push_back({ "MPC_MIXED_COST_WITH_BOUND_CONSTRAINTS", [] { boundedCase<copra::TargetCost>(true); } });
        groups.push_back({ "MPC_TARGET_COST_WITH_INEQUALITY_CONSTRAINTS", ineqCase<copra::TargetCost> });
        groups.push_back({ "MPC_TRAJECTORY_COST_WITH_INEQUALITY_CONSTRAINTS", ineqCase<copra::TrajectoryCost> });
        groups.push_back({ "MPC_TARGET_COST_WITH_MIXED_CONSTRAINTS", mixedCase<copra::TargetCost> });
        groups.push_back({ "MPC_TRAJECTORY_COST_WITH_MIXED_CONSTRAINTS", mixedCase<copra::TrajectoryCost> });
        groups.push_back({ "MPC_TARGET_COST_WITH_EQUALITY_CONSTRAINTS", eqCase<copra::TargetCost> });
        groups.push_back({ "MPC_TRAJECTORY_COST_WITH_EQUALITY_CONSTRAINTS", eqCase<copra::TrajectoryCost> });
        groups.push_back({ "REMOVE_COST_AND_CONSTRAINT", removeAndDelete });
        groups.push_back({ "PLUGIN_PROTOCOL_AND_GETTERS", pluginProtocolAndGetters });
        groups.push_back({ "LMPC_AND_INITIAL-STATE-LMPC_COMPARISON", initialStateComparison });
        groups.push_back({ "FULL_SIZE_ENTRIES_SOLVE", fullSizeEntriesSolve });
        groups.push_back({ "BATCHED_ENTRY", batchedEntry });
        groups.push_back({ "TWO_HOST_THREADS", twoHostThreads });
    }
    for (auto& g : groups) {
        const int before = g_fail;
        try {
            g.second();
        } catch (const std::exception& e) {
            ++g_fail;
            std::printf("  EXCEPTION in %s: %s\n", g.first.c_str(), e.what());
        }
        std::printf("[%s] %s\n", g_fail == before ? " ok " : "FAIL", g.first.c_str());
    }
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail == 0 ? 0 : 1;
}
