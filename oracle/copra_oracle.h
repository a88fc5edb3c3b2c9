/*
 * copra_oracle.h -- C interface of the CPU ORACLE (test infrastructure, NOT the product).
 *
 * The oracle is a plain C++14 restatement (no Eigen, scalar loops) of the reference hot path
 *   PreviewSystem::updateSystem -> cost/constraint update -> LMPC::makeQPForm ->
 *   QuadProgDenseSolver::SI_solve -> Eigen::QuadProgDense (qpgen2) -> LMPC::updateResults
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (libcopra_b200.so) never links or calls it.
 *
 * All matrices are column-major with ld == rows (Eigen default), all reals are IEEE float64.
 */
#ifndef COPRA_ORACLE_H
#define COPRA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* cost kinds: reference include/costFunctions.h:103-219 */
enum { ORC_COST_TRAJECTORY = 0, ORC_COST_TARGET = 1, ORC_COST_CONTROL = 2, ORC_COST_MIXED = 3 };
/* constraint kinds: reference include/constraints.h:114-307 */
enum {
    ORC_CSTR_TRAJECTORY = 0,
    ORC_CSTR_CONTROL = 1,
    ORC_CSTR_MIXED = 2,
    ORC_CSTR_TRAJECTORY_BOUND = 3,
    ORC_CSTR_CONTROL_BOUND = 4
};

/* One cost function as the user hands it to copra (before initializeCost).
 *   TRAJECTORY / TARGET: M (rows x colsM), p, w           (N unused)
 *   CONTROL            : N (rows x colsN), p, w           (M unused)
 *   MIXED              : M (rows x colsM), N (rows x colsN), p, w
 * colsM == nx (step-size) or nx*(Nsteps+1) (full-size); colsN == nu or nu*Nsteps.
 * w has `wrows` entries and is expanded like CostFunction::weights() (costFunctions.h:54-67). */
typedef struct {
    int kind;
    int rows;
    int colsM;
    int colsN;
    int wrows;
    int autospan; /* call autoSpan() before initializeCost (reference AutoSpan.cpp) */
    const double* M;
    const double* N;
    const double* p;
    const double* w;
} orc_cost;

/* One constraint.
 *   TRAJECTORY      : E (rows x colsE), f
 *   CONTROL         : G (rows x colsG), f
 *   MIXED           : E, G, f
 *   TRAJECTORY_BOUND: lower, upper (rows entries each; +-inf lines are skipped)
 *   CONTROL_BOUND   : lower, upper (rows entries each) */
typedef struct {
    int kind;
    int rows;
    int colsE;
    int colsG;
    int is_ineq;
    int autospan;
    const double* E;
    const double* G;
    const double* f;
    const double* lower;
    const double* upper;
} orc_constraint;

typedef struct {
    int nx, nu, N;
    const double *A, *B, *d, *x0;
    int ncost;
    const orc_cost* costs;
    int ncstr;
    const orc_constraint* cstrs;
    /* InitialStateLMPC (reference src/InitialStateLMPC.cpp): decision vector [x0; U] */
    int initial_state;
    const double *R, *r, *x0lb, *x0ub; /* may be NULL -> reference defaults (R=0,r=0,bounds=x0) */
} orc_problem;

/* Sizes implied by a problem (so callers can allocate). */
typedef struct {
    int X;    /* nx*(N+1)  */
    int nU;   /* nu*N      */
    int nvar; /* nU or nx+nU */
    int meq, mineq;
} orc_sizes;

/* Every output pointer may be NULL. */
typedef struct {
    double *Phi, *Psi, *xi;               /* X x nx, X x nU, X */
    double *Q, *c;                        /* nvar x nvar, nvar */
    double *Aeq, *beq, *Aineq, *bineq;    /* meq x nvar, meq, mineq x nvar, mineq */
    double *lb, *ub;                      /* nvar */
    double* x;                            /* nvar : raw QP result */
    double* control;                      /* nU */
    double* trajectory;                   /* X */
    int* iact;                            /* meq+mineq+2*nvar, 1-based, add order, 0 padded */
    int* nact;
    int* iter;                            /* 2 */
    int* fail;                            /* 0 ok / 1 infeasible / 2 not PD */
    double* lagr;                         /* meq+mineq+2*nvar multipliers */
    double* crval;                        /* objective value */
    double* t_build;                      /* seconds: everything but SI_solve */
    double* t_solve;                      /* seconds: SI_solve only (LMPC.cpp:90-92) */
} orc_outputs;

/* returns 0 on success, -1 std::domain_error, -2 std::runtime_error (message via orc_last_error) */
int orc_sizes_of(const orc_problem* p, orc_sizes* s);
int orc_condense(int nx, int nu, int N, const double* A, const double* B, const double* d,
    double* Phi, double* Psi, double* xi);
/* build the QP (no solve if out->x == NULL && out->control == NULL) and optionally solve */
int orc_lmpc(const orc_problem* p, orc_outputs* out);

/* QuadProgDenseSolver::SI_solve + Eigen::QuadProgDense::solve + qpgen2, on a raw QP:
 *   min 1/2 x'Qx + c'x  s.t. Aeq x = beq, Aineq x <= bineq, lb <= x <= ub
 * iact: 1-based indices in [eq | ineq | upper | lower] space (q = meq+m+2n entries).
 * returns fail code (0/1/2). */
int orc_quadprog(int n, int meq, int m, const double* Q, const double* c, const double* Aeq,
    const double* beq, const double* Aineq, const double* bineq, const double* lb,
    const double* ub, double* x, int* iact, int* nact, int* iter, double* lagr, double* crval);

/* CPU baseline: solve `batch` problems (array of orc_problem), one instance per thread, `threads`
 * workers.  Writes control (nU x batch), trajectory (X x batch), fail (batch), iter (2 x batch),
 * per-instance total seconds `t_inst` (batch).  Returns wall seconds, <0 on error. */
double orc_lmpc_batch(const orc_problem* probs, int batch, int threads, double* control,
    double* trajectory, int* fail, int* iter, int* nact, int* iact, double* t_inst);

int orc_hw_threads(void);
const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
