/* jbugs_gq.h -- generated quantities for all draws of a run at once, on the device.
 *
 * Replaces, for S draws per launch instead of one model evaluation per draw on the host:
 *   /root/reference/JuliaBUGS/src/model/abstractmcmc.jl:92-110   gen_chains: re-evaluates the model for every draw to
 *                                                                 fill in the nodes that are not part of the target
 *   /root/reference/JuliaBUGS/src/model/evaluation.jl:354-378     evaluate_with_rng!!: deterministic nodes from their
 *                                                                 node functions, unobserved stochastic nodes without
 *                                                                 observed descendants by a draw from their distribution
 *
 * A program is a list of statements in dependency order.  A statement covers the cells of one BUGS statement's loop
 * domain; it owns `n_cells` consecutive rows of the output matrix Q[n_rows][ld_out] (one column per draw) and a short
 * SSA tape that is evaluated once per (cell, draw).  A logical statement stores the tape's last value; a stochastic one
 * draws from `family` with the tape values in `arg_slot[]` as arguments (Philox4x32-10, counter = (row, draw), key =
 * seed: reproducible, independent of the launch geometry).
 *
 * Tape leaves: a constant; the CONSTRAINED value of a parameter (theta row, pushed through its bijector:
 * evaluation.jl:243-251); an earlier row of Q; a per-cell data value.  Row numbers and data values that vary over the
 * loop domain come from per-cell arrays (`ipool`, `dpool`) computed by the host from the statement's index expressions.
 */
#ifndef JBUGS_GQ_H
#define JBUGS_GQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JBUGS_GQ_MAX_OPS 96
#define JBUGS_GQ_LOGICAL 0xFFFFFFFFu /* jbugs_gq_stmt.family of a deterministic statement */

enum jbugs_gq_leaf {
    JBUGS_GQ_CONST = 0,      /* value */
    JBUGS_GQ_THETA = 1,      /* constrained value of theta row `offset` */
    JBUGS_GQ_THETA_CELL = 2, /* ... of theta row ipool[offset + cell] */
    JBUGS_GQ_Q = 3,          /* Q row `offset` */
    JBUGS_GQ_Q_CELL = 4,     /* Q row ipool[offset + cell] */
    JBUGS_GQ_DATA_CELL = 5   /* dpool[offset + cell] */
};

/* ops: the values of enum `jbugs_op` in jbugs_plan.h -- 0..7 unary, 16..20 binary, 32 leaf -- forward values only -- plus: */
enum jbugs_gq_op {
    JBUGS_GQ_OP_ABS = 40, JBUGS_GQ_OP_STEP = 41, JBUGS_GQ_OP_LOGIT = 42, JBUGS_GQ_OP_CLOGLOG = 43, JBUGS_GQ_OP_PROBIT = 44,
    JBUGS_GQ_OP_SIN = 45, JBUGS_GQ_OP_COS = 46, JBUGS_GQ_OP_ROUND = 47, JBUGS_GQ_OP_TRUNC = 48, JBUGS_GQ_OP_LOGGAM = 49,
    JBUGS_GQ_OP_LOGFACT = 50,
    JBUGS_GQ_OP_EQUALS = 60, JBUGS_GQ_OP_MAX = 61, JBUGS_GQ_OP_MIN = 62 /* binary */
};

typedef struct jbugs_gq_op_rec {
    uint32_t op;
    uint16_t a, b;      /* operand slots (earlier ops of the same statement) */
    uint32_t leaf_kind; /* op == JBUGS_OP_LEAF */
    int64_t offset;
    double value;
} jbugs_gq_op_rec; /* 32 bytes */

typedef struct jbugs_gq_stmt {
    int64_t n_cells;
    int64_t out_row0;  /* the statement writes Q rows out_row0 .. out_row0 + n_cells - 1 */
    uint32_t first_op, n_ops;
    uint32_t family;   /* enum jbugs_family, or JBUGS_GQ_LOGICAL */
    uint32_t n_args;
    uint32_t arg_slot[3];
    uint32_t pad;
} jbugs_gq_stmt; /* 48 bytes */

typedef struct jbugs_gq jbugs_gq;

/* `transform`, `lb`, `ub`: bijector of every theta row (enum jbugs_transform and its bounds), length D. */
int jbugs_gq_create(int device, int64_t D, const int32_t* transform, const double* lb, const double* ub,
                    const jbugs_gq_stmt* stmts, int64_t n_stmts, const jbugs_gq_op_rec* ops, int64_t n_ops,
                    const int64_t* ipool, int64_t n_ipool, const double* dpool, int64_t n_dpool, int64_t n_rows,
                    jbugs_gq** out);
void jbugs_gq_destroy(jbugs_gq* gq);

/* theta: R x S unconstrained draws, chain-fast (theta[r * ld + s]); `rows[r]` is the model's theta index of input row r
 * (NULL: R == D, identity).  out: n_rows x S (out[row * ld_out + s]).  Host or device pointers (all of one kind).
 * A program that reads a theta row the caller did not supply fails with JBUGS_ERR_INVALID. */
int jbugs_gq_run(jbugs_gq* gq, const double* theta, int64_t ld, int64_t R, const int64_t* rows, int64_t S, uint64_t seed,
                 double* out, int64_t ld_out, void* stream);
int64_t jbugs_gq_launch_count(const jbugs_gq* gq); /* kernels launched so far */

#ifdef __cplusplus
}
#endif
#endif
