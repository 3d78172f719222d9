"""Example models in BUGS syntax (classic BUGS Vol. 1), written out here as model text.

The numeric fixtures (data, inits) live in `tests/golden/examples_vol1.json`, extracted from the
reference tree by `tests/golden/make_fixtures.py`.  The reference ships the same models as `@bugs`
expressions in `JuliaBUGS/src/BUGSExamples/Volume_1/` (`01_Rats.jl:3-19`, `02_Pumps.jl:3-11`,
`03_Dogs.jl:3-21`, `04_Seeds.jl:3-16`, `05_Surgical.jl:4-9,31-41`, `11_Epil.jl:3-45`,
`12_Blocker.jl:3-16`, `15_Bones.jl:3-22`).
"""

RATS = """
model {
    for (i in 1:N) {
        for (j in 1:T) {
            Y[i, j] ~ dnorm(mu[i, j], tau.c)
            mu[i, j] <- alpha[i] + beta[i] * (x[j] - xbar)
        }
        alpha[i] ~ dnorm(alpha.c, alpha.tau)
        beta[i] ~ dnorm(beta.c, beta.tau)
    }
    tau.c ~ dgamma(0.001, 0.001)
    sigma <- 1 / sqrt(tau.c)
    alpha.c ~ dnorm(0.0, 1.0E-6)
    alpha.tau ~ dgamma(0.001, 0.001)
    beta.c ~ dnorm(0.0, 1.0E-6)
    beta.tau ~ dgamma(0.001, 0.001)
    alpha0 <- alpha.c - xbar * beta.c
}
"""

PUMPS = """
model {
    for (i in 1:N) {
        theta[i] ~ dgamma(alpha, beta)
        lambda[i] <- theta[i] * t[i]
        x[i] ~ dpois(lambda[i])
    }
    alpha ~ dexp(1)
    beta ~ dgamma(0.1, 1.0)
}
"""

DOGS = """
model {
    for (i in 1:Dogs) {
        xa[i, 1] <- 0
        xs[i, 1] <- 0
        p[i, 1] <- 0
        for (j in 2:Trials) {
            xa[i, j] <- sum(Y[i, 1:(j - 1)])
            xs[i, j] <- j - 1 - xa[i, j]
            log(p[i, j]) <- alpha * xa[i, j] + beta * xs[i, j]
            y[i, j] <- 1 - Y[i, j]
            y[i, j] ~ dbern(p[i, j])
        }
    }
    alpha ~ dunif(-10, -0.00001)
    beta ~ dunif(-10, -0.00001)
    A <- exp(alpha)
    B <- exp(beta)
}
"""

SEEDS = """
model {
    for (i in 1:N) {
        r[i] ~ dbin(p[i], n[i])
        b[i] ~ dnorm(0.0, tau)
        logit(p[i]) <- alpha0 + alpha1 * x1[i] + alpha2 * x2[i] + alpha12 * x1[i] * x2[i] + b[i]
    }
    alpha0 ~ dnorm(0.0, 1.0E-6)
    alpha1 ~ dnorm(0.0, 1.0E-6)
    alpha2 ~ dnorm(0.0, 1.0E-6)
    alpha12 ~ dnorm(0.0, 1.0E-6)
    tau ~ dgamma(0.001, 0.001)
    sigma <- 1 / sqrt(tau)
}
"""

SURGICAL_SIMPLE = """
model {
    for (i in 1:N) {
        p[i] ~ dbeta(1.0, 1.0)
        r[i] ~ dbin(p[i], n[i])
    }
}
"""

SURGICAL_REALISTIC = """
model {
    for (i in 1:N) {
        b[i] ~ dnorm(mu, tau)
        r[i] ~ dbin(p[i], n[i])
        logit(p[i]) <- b[i]
    }
    pop.mean <- exp(mu) / (1 + exp(mu))
    mu ~ dnorm(0.0, 1.0E-6)
    sigma <- 1 / sqrt(tau)
    tau ~ dgamma(0.001, 0.001)
}
"""

EPIL = """
model {
    for (j in 1:N) {
        for (k in 1:T) {
            log(mu[j, k]) <- a0 + alpha.Base * (log.Base4[j] - log.Base4.bar)
                + alpha.Trt * (Trt[j] - Trt.bar)
                + alpha.BT * (BT[j] - BT.bar)
                + alpha.Age * (log.Age[j] - log.Age.bar)
                + alpha.V4 * (V4[k] - V4.bar) + b1[j] + b[j, k]
            y[j, k] ~ dpois(mu[j, k])
            b[j, k] ~ dnorm(0.0, tau.b)
        }
        b1[j] ~ dnorm(0.0, tau.b1)
        BT[j] <- Trt[j] * log.Base4[j]
        log.Base4[j] <- log(Base[j] / 4)
        log.Age[j] <- log(Age[j])
    }
    log.Age.bar <- mean(log.Age[])
    Trt.bar <- mean(Trt[])
    BT.bar <- mean(BT[])
    log.Base4.bar <- mean(log.Base4[])
    V4.bar <- mean(V4[])
    a0 ~ dnorm(0.0, 1.0E-4)
    alpha.Base ~ dnorm(0.0, 1.0E-4)
    alpha.Trt ~ dnorm(0.0, 1.0E-4)
    alpha.BT ~ dnorm(0.0, 1.0E-4)
    alpha.Age ~ dnorm(0.0, 1.0E-4)
    alpha.V4 ~ dnorm(0.0, 1.0E-4)
    tau.b1 ~ dgamma(1.0E-3, 1.0E-3)
    sigma.b1 <- 1.0 / sqrt(tau.b1)
    tau.b ~ dgamma(1.0E-3, 1.0E-3)
    sigma.b <- 1.0 / sqrt(tau.b)
    alpha0 <- a0 - alpha.Base * log.Base4.bar - alpha.Trt * Trt.bar - alpha.BT * BT.bar
        - alpha.Age * log.Age.bar - alpha.V4 * V4.bar
}
"""

BLOCKERS = """
model {
    for (i in 1:Num) {
        rc[i] ~ dbin(pc[i], nc[i])
        rt[i] ~ dbin(pt[i], nt[i])
        logit(pc[i]) <- mu[i]
        logit(pt[i]) <- mu[i] + delta[i]
        mu[i] ~ dnorm(0.0, 1.0E-5)
        delta[i] ~ dnorm(d, tau)
    }
    d ~ dnorm(0.0, 1.0E-6)
    tau ~ dgamma(0.001, 0.001)
    delta.new ~ dnorm(d, tau)
    sigma <- 1 / sqrt(tau)
}
"""

BONES = """
model {
    for (i in 1:nChild) {
        theta[i] ~ dnorm(0.0, 0.001)
        for (j in 1:nInd) {
            for (k in 1:ncat[j] - 1) {
                logit(Q[i, j, k]) <- delta[j] * (theta[i] - gamma[j, k])
            }
        }
        for (j in 1:nInd) {
            p[i, j, 1] <- 1 - Q[i, j, 1]
            for (k in 2:ncat[j] - 1) {
                p[i, j, k] <- Q[i, j, k - 1] - Q[i, j, k]
            }
            p[i, j, ncat[j]] <- Q[i, j, ncat[j] - 1]
            grade[i, j] ~ dcat(p[i, j, 1:ncat[j]])
        }
    }
}
"""

# Dense hierarchical logistic regression (BASELINE.json config 4): inprod -> GEMM path
DENSE_LOGISTIC = """
model {
    for (i in 1:N) {
        y[i] ~ dbern(p[i])
        logit(p[i]) <- inprod(X[i, 1:P], beta[1:P])
    }
    for (j in 1:P) {
        beta[j] ~ dnorm(mu.b, tau.b)
    }
    mu.b ~ dnorm(0.0, 1.0E-6)
    tau.b ~ dgamma(0.001, 0.001)
}
"""

# Salm: extra-Poisson variation in a dose-response study (BUGS examples vol. 1)
SALM = """
model {
    for (i in 1:doses) {
        for (j in 1:plates) {
            y[i, j] ~ dpois(mu[i, j])
            log(mu[i, j]) <- alpha + beta * log(x[i] + 10) + gamma * x[i] + lambda[i, j]
            lambda[i, j] ~ dnorm(0.0, tau)
        }
    }
    alpha ~ dnorm(0.0, 1.0E-6)
    beta ~ dnorm(0.0, 1.0E-6)
    gamma ~ dnorm(0.0, 1.0E-6)
    tau ~ dgamma(0.001, 0.001)
    sigma <- 1 / sqrt(tau)
}
"""

# Equiv: bioequivalence in a cross-over trial; the subject effect is indexed by the INNER loop variable
EQUIV = """
model {
    for (k in 1:P) {
        for (i in 1:N) {
            Y[i, k] ~ dnorm(m[i, k], tau1)
            m[i, k] <- mu + sign[T[i, k]] * phi / 2 + sign[k] * pi / 2 + delta[i]
            T[i, k] <- group[i] * (k - 1.5) + 1.5
        }
    }
    for (i in 1:N) {
        delta[i] ~ dnorm(0.0, tau2)
    }
    tau1 ~ dgamma(0.001, 0.001)
    sigma1 <- 1 / sqrt(tau1)
    tau2 ~ dgamma(0.001, 0.001)
    sigma2 <- 1 / sqrt(tau2)
    mu ~ dnorm(0.0, 1.0E-6)
    phi ~ dnorm(0.0, 1.0E-6)
    pi ~ dnorm(0.0, 1.0E-6)
    theta <- exp(phi)
    equiv <- step(theta - 0.8) - step(theta - 1.2)
}
"""

# Dyes: variance components
DYES = """
model {
    for (i in 1:batches) {
        mu[i] ~ dnorm(theta, tau.btw)
        for (j in 1:samples) {
            y[i, j] ~ dnorm(mu[i], tau.with)
        }
    }
    sigma2.with <- 1 / tau.with
    sigma2.btw <- 1 / tau.btw
    tau.with ~ dgamma(0.001, 0.001)
    tau.btw ~ dgamma(0.001, 0.001)
    theta ~ dnorm(0.0, 1.0E-10)
}
"""

# Oxford: smooth fit to log-odds ratios; two binomial statements sharing mu[i]
OXFORD = """
model {
    for (i in 1:K) {
        r0[i] ~ dbin(p0[i], n0[i])
        r1[i] ~ dbin(p1[i], n1[i])
        logit(p0[i]) <- mu[i]
        logit(p1[i]) <- mu[i] + logPsi[i]
        logPsi[i] <- alpha + beta1 * year[i] + beta2 * (year[i] * year[i] - 22) + b[i]
        b[i] ~ dnorm(0, tau)
        mu[i] ~ dnorm(0.0, 1.0E-6)
    }
    alpha ~ dnorm(0.0, 1.0E-6)
    beta1 ~ dnorm(0.0, 1.0E-6)
    beta2 ~ dnorm(0.0, 1.0E-6)
    tau ~ dgamma(0.001, 0.001)
    sigma <- 1 / sqrt(tau)
}
"""

# Stacks: regression on standardised covariates; the coefficient vector is referenced element by element
STACKS = """
model {
    for (j in 1:p) {
        b[j] <- beta[j] / sd(x[, j])
        for (i in 1:N) {
            z[i, j] <- (x[i, j] - mean(x[, j])) / sd(x[, j])
        }
    }
    b0 <- beta0 - b[1] * mean(x[, 1]) - b[2] * mean(x[, 2]) - b[3] * mean(x[, 3])
    for (i in 1:N) {
        Y[i] ~ dnorm(mu[i], tau)
        mu[i] <- beta0 + beta[1] * z[i, 1] + beta[2] * z[i, 2] + beta[3] * z[i, 3]
        stres[i] <- (Y[i] - mu[i]) / sigma
        outlier[i] <- step(stres[i] - 2.5) + step(-(stres[i] + 2.5))
    }
    beta0 ~ dnorm(0, 0.00001)
    for (j in 1:p) {
        beta[j] ~ dnorm(0, 0.00001)
    }
    tau ~ dgamma(1.0E-3, 1.0E-3)
    sigma <- sqrt(1 / tau)
}
"""

# Beetles (BUGS examples vol. 2): dose-response with a choice of link function; BEETLES % "logit" | "probit" | "cloglog"
BEETLES = """
model {
    for (i in 1:N) {
        r[i] ~ dbin(p[i], n[i])
        %s(p[i]) <- alpha.star + beta * (x[i] - mean(x[]))
        rhat[i] <- n[i] * p[i]
    }
    alpha <- alpha.star - beta * mean(x[])
    beta ~ dnorm(0.0, 0.001)
    alpha.star ~ dnorm(0.0, 0.001)
}
"""
BEETLES_DATA = {"x": [1.6907, 1.7242, 1.7552, 1.7842, 1.8113, 1.8369, 1.8610, 1.8839],
                "n": [59, 60, 62, 56, 63, 59, 62, 60], "r": [6, 13, 18, 28, 52, 53, 61, 60], "N": 8}

# ---- models whose predictor is not linear in the parameters: lowered to expression plates ---------------------------
# Dugongs (BUGS examples vol. 2): non-linear growth curve, no random effects
DUGONGS = """
model {
    for (i in 1:N) {
        Y[i] ~ dnorm(mu[i], tau)
        mu[i] <- alpha - beta * pow(gamma, x[i])
    }
    alpha ~ dunif(0, 100)
    beta ~ dunif(0, 100)
    gamma ~ dunif(0.5, 1.0)
    tau ~ dgamma(0.001, 0.001)
    sigma <- 1 / sqrt(tau)
    U3 <- logit(gamma)
}
"""

# LSAT (vol. 1): Rasch model; the response patterns are expanded by a ragged data transformation
LSAT = """
model {
    for (j in 1:culm[1]) {
        for (k in 1:T) {
            r[j, k] <- response[1, k]
        }
    }
    for (i in 2:R) {
        for (j in culm[i - 1] + 1:culm[i]) {
            for (k in 1:T) {
                r[j, k] <- response[i, k]
            }
        }
    }
    for (j in 1:N) {
        for (k in 1:T) {
            logit(p[j, k]) <- beta * theta[j] - alpha[k]
            r[j, k] ~ dbern(p[j, k])
        }
        theta[j] ~ dnorm(0, 1)
    }
    for (k in 1:T) {
        alpha[k] ~ dnorm(0, 0.0001)
        a[k] <- alpha[k] - mean(alpha[])
    }
    beta ~ dunif(0, 1000)
}
"""

# Air (vol. 2): measurement error in the covariate of a logistic regression -- a latent covariate X[j] multiplies a
# coefficient
AIR = """
model {
    for (j in 1:J) {
        y[j] ~ dbin(p[j], n[j])
        logit(p[j]) <- theta[1] + theta[2] * X[j]
        X[j] ~ dnorm(mu[j], tau)
        mu[j] <- alpha + beta * Z[j]
    }
    theta[1] ~ dnorm(0.0, 0.001)
    theta[2] ~ dnorm(0.0, 0.001)
}
"""

# Leuk (vol. 1): Cox regression as a Poisson counting process with independent-increment gamma hazards
LEUK = """
model {
    for (i in 1:N) {
        for (j in 1:T) {
            Y[i, j] <- step(obs.t[i] - t[j] + eps)
            dN[i, j] <- Y[i, j] * step(t[j + 1] - obs.t[i] - eps) * fail[i]
        }
    }
    for (j in 1:T) {
        for (i in 1:N) {
            dN[i, j] ~ dpois(Idt[i, j])
            Idt[i, j] <- Y[i, j] * exp(beta * Z[i]) * dL0[j]
        }
        dL0[j] ~ dgamma(mu[j], c)
        mu[j] <- dL0.star[j] * c
        S.treat[j] <- pow(exp(-sum(dL0[1:j])), exp(beta * -0.5))
        S.placebo[j] <- pow(exp(-sum(dL0[1:j])), exp(beta * 0.5))
    }
    c <- 0.001
    r <- 0.1
    for (j in 1:T) {
        dL0.star[j] <- r * (t[j + 1] - t[j])
    }
    beta ~ dnorm(0.0, 0.000001)
}
"""

# LeukFr: the Cox model of Leuk with a frailty per matched pair, b[pair[i]] -- a random effect picked by a data index of
# the inner loop variable inside a predictor that is not linear in the parameters
LEUKFR = """
model {
    for (i in 1:N) {
        for (j in 1:T) {
            Y[i, j] <- step(obs.t[i] - t[j] + eps)
            dN[i, j] <- Y[i, j] * step(t[j + 1] - obs.t[i] - eps) * fail[i]
        }
    }
    for (j in 1:T) {
        for (i in 1:N) {
            dN[i, j] ~ dpois(Idt[i, j])
            Idt[i, j] <- Y[i, j] * exp(beta * Z[i] + b[pair[i]]) * dL0[j]
        }
        dL0[j] ~ dgamma(mu[j], c)
        mu[j] <- dL0.star[j] * c
        S.treat[j] <- pow(exp(-sum(dL0[1:j])), exp(beta * -0.5))
        S.placebo[j] <- pow(exp(-sum(dL0[1:j])), exp(beta * 0.5))
    }
    for (k in 1:Npairs) {
        b[k] ~ dnorm(0.0, tau)
    }
    tau ~ dgamma(0.001, 0.001)
    sigma <- sqrt(1 / tau)
    c <- 0.001
    r <- 0.1
    for (j in 1:T) {
        dL0.star[j] <- r * (t[j + 1] - t[j])
    }
    beta ~ dnorm(0.0, 0.000001)
}
"""

MICE = """
model {
    for (i in 1:M) {
        for (j in 1:N) {
            t[i, j] ~ dweib(r, mu[i]) C(t.cen[i, j],)
        }
        mu[i] <- exp(beta[i])
        beta[i] ~ dnorm(0.0, 0.001)
        median[i] <- pow(log(2) * exp(-beta[i]), 1 / r)
    }
    r ~ dunif(0.1, 10)
    veh.control <- beta[2] - beta[1]
    test.sub <- beta[3] - beta[1]
    pos.control <- beta[4] - beta[1]
}
"""

KIDNEY = """
model {
    for (i in 1:N) {
        for (j in 1:M) {
            t[i, j] ~ dweib(r, mu[i, j]) C(t.cen[i, j],)
            log(mu[i, j]) <- alpha + beta.age * age[i, j] + beta.sex * sex[i] + beta.dis[disease[i]] + b[i]
        }
        b[i] ~ dnorm(0.0, tau)
    }
    alpha ~ dnorm(0.0, 0.0001)
    beta.age ~ dnorm(0.0, 0.0001)
    beta.sex ~ dnorm(0.0, 0.0001)
    for (k in 2:4) {
        beta.dis[k] ~ dnorm(0.0, 0.0001)
    }
    tau ~ dgamma(1.0E-3, 1.0E-3)
    r ~ dgamma(1.0, 1.0E-3)
    sigma <- 1 / sqrt(tau)
}
"""

ORANGE = """
model {
    for (i in 1:K) {
        for (j in 1:n) {
            Y[i, j] ~ dnorm(eta[i, j], tauC)
            eta[i, j] <- phi[i, 1] / (1 + phi[i, 2] * exp(phi[i, 3] * x[j]))
        }
        phi[i, 1] <- exp(theta[i, 1])
        phi[i, 2] <- exp(theta[i, 2]) - 1
        phi[i, 3] <- -exp(theta[i, 3])
        for (k in 1:3) {
            theta[i, k] ~ dnorm(mu[k], tau[k])
        }
    }
    tauC ~ dgamma(1.0E-3, 1.0E-3)
    sigma.C <- 1 / sqrt(tauC)
    for (k in 1:3) {
        mu[k] ~ dnorm(0, 1.0E-4)
        tau[k] ~ dgamma(1.0E-3, 1.0E-3)
        sigma[k] <- 1 / sqrt(tau[k])
    }
}
"""

# vol. 2, Cervix: a case-control study with errors in covariates.  x[i] (true exposure) is observed for the first 115
# subjects only; the missing x[i] are finite discrete nodes -- parameters with the identity bijector in the reference's
# UseGraph / generated modes (outside the plate class), summed out under UseAutoMarginalization
CERVIX = """
model {
    for (i in 1:N) {
        x[i] ~ dbern(q)
        logit(p[i]) <- beta0C + beta * x[i]
        d[i] ~ dbern(p[i])
        x1[i] <- x[i] + 1
        d1[i] <- d[i] + 1
        w[i] ~ dbern(phi[x1[i], d1[i]])
    }
    q ~ dunif(0.0, 1.0)
    beta0C ~ dnorm(0.0, 0.00001)
    beta ~ dnorm(0.0, 0.00001)
    for (j in 1:2) {
        for (k in 1:2) {
            phi[j, k] ~ dunif(0.0, 1.0)
        }
    }
    gamma1 <- 1 / (1 + (1 + exp(beta0C + beta)) / (1 + exp(beta0C)) * (1 - q) / q)
    gamma2 <- 1 / (1 + (1 + exp(-beta0C - beta)) / (1 + exp(-beta0C)) * (1 - q) / q)
}
"""

# vol. 1, Magnesium: one meta-analysis under six priors for the between-study variance.  The `rtx[j, k] <- rt[k]` lines
# are the double-definition idiom (the data transformation makes rtx / rcx observed); prior 6 is a half-normal, T(0, )
MAGNESIUM = """
model {
    for (j in 1:6) {
        mu[j] ~ dunif(-10, 10)
        odds.ratio[j] <- exp(mu[j])
        for (k in 1:8) {
            theta[j, k] ~ dnorm(mu[j], inv.tau.sqrd[j])
            rtx[j, k] ~ dbin(pt[j, k], nt[k])
            rtx[j, k] <- rt[k]
            rcx[j, k] ~ dbin(pc[j, k], nc[k])
            rcx[j, k] <- rc[k]
            logit(pt[j, k]) <- theta[j, k] + phi[j, k]
            phi[j, k] <- logit(pc[j, k])
            pc[j, k] ~ dunif(0, 1)
        }
    }
    for (k in 1:8) {
        y[k] <- log(((rt[k] + 0.5) / (nt[k] - rt[k] + 0.5)) / ((rc[k] + 0.5) / (nc[k] - rc[k] + 0.5)))
        sigma.sqrd[k] <- 1 / (rt[k] + 0.5) + 1 / (nt[k] - rt[k] + 0.5) + 1 / (rc[k] + 0.5) + 1 / (nc[k] - rc[k] + 0.5)
        prec.sqrd[k] <- 1 / sigma.sqrd[k]
    }
    s0.sqrd <- 1 / mean(prec.sqrd[1:8])
    inv.tau.sqrd[1] ~ dgamma(0.001, 0.001)
    tau.sqrd[1] <- 1 / inv.tau.sqrd[1]
    tau[1] <- sqrt(tau.sqrd[1])
    tau.sqrd[2] ~ dunif(0, 50)
    tau[2] <- sqrt(tau.sqrd[2])
    inv.tau.sqrd[2] <- 1 / tau.sqrd[2]
    tau[3] ~ dunif(0, 50)
    tau.sqrd[3] <- tau[3] * tau[3]
    inv.tau.sqrd[3] <- 1 / tau.sqrd[3]
    B0 ~ dunif(0, 1)
    tau.sqrd[4] <- s0.sqrd * (1 - B0) / B0
    tau[4] <- sqrt(tau.sqrd[4])
    inv.tau.sqrd[4] <- 1 / tau.sqrd[4]
    D0 ~ dunif(0, 1)
    tau[5] <- sqrt(s0.sqrd) * (1 - D0) / D0
    tau.sqrd[5] <- tau[5] * tau[5]
    inv.tau.sqrd[5] <- 1 / tau.sqrd[5]
    p0 <- phi(0.75) / s0.sqrd
    tau.sqrd[6] ~ dnorm(0, p0) T(0, )
    tau[6] <- sqrt(tau.sqrd[6])
    inv.tau.sqrd[6] <- 1 / tau.sqrd[6]
}
"""

MODELS = {
    "leukfr": LEUKFR,
    "magnesium": MAGNESIUM,
    "cervix": CERVIX,
    "orange": ORANGE,
    "mice": MICE,
    "kidney": KIDNEY,
    "dugongs": DUGONGS,
    "lsat": LSAT,
    "air": AIR,
    "leuk": LEUK,
    "stacks": STACKS,
    "salm": SALM,
    "equiv": EQUIV,
    "dyes": DYES,
    "oxford": OXFORD,
    "rats": RATS,
    "pumps": PUMPS,
    "dogs": DOGS,
    "seeds": SEEDS,
    "surgical_simple": SURGICAL_SIMPLE,
    "surgical_realistic": SURGICAL_REALISTIC,
    "epil": EPIL,
    "blockers": BLOCKERS,
    "bones": BONES,
    "dense_logistic": DENSE_LOGISTIC,
}
