# ldsr_b200.R -- drop-in R wrappers that route ldsr's EM fan-out through ONE batched GPU call.
#
# NOT EXERCISED IN THIS REPOSITORY (no R in the build image); the Python mirror ldsr_b200/api.py
# implements the same logic and is what the tests run.  Source these definitions after
# library(ldsr) (or paste them over the originals in R/LDS_reconstruction.R) -- names, arguments
# and return values are the reference's.
#
# Random numbers: make_init() is called here, on the R side, in the order the reference's
# workers would call it under registerDoSEQ() (fold-major, then ensemble member), so set.seed()
# gives the same initial values -- and hence the same selected restarts -- as the reference.

.theta_vec <- function(th) c(th$A, th$B, th$C, th$D, th$Q, th$R, th$mu1, th$V1)

.theta_list <- function(vec, p, q) {
  th <- list(A = matrix(vec[1]), B = matrix(vec[2:(p + 1)], 1, p), C = matrix(vec[p + 2]),
             D = matrix(vec[(p + 3):(p + q + 2)], 1, q), Q = matrix(vec[p + q + 3]),
             R = matrix(vec[p + q + 4]), mu1 = matrix(vec[p + q + 5]), V1 = matrix(vec[p + q + 6]))
  class(th) <- 'theta'
  th
}

# One batched call: `jobs` is a list of list(series = index, held = 1-based steps, init = list of theta)
.em_batch <- function(series, jobs, niter, tol, n.devices = 0L) {
  stride <- max(sapply(series, function(s) nrow(s$u) + nrow(s$v) + 6L))
  theta0 <- do.call(cbind, lapply(jobs, function(j) sapply(j$init, function(th) {
    v <- .theta_vec(th); c(v, rep(0, stride - length(v)))
  })))
  .Call('_ldsr_em_batch', series, as.integer(sapply(jobs, `[[`, 'series')),
        lapply(jobs, function(j) as.integer(j$held)),
        rep(seq_along(jobs), sapply(jobs, function(j) length(j$init))),
        theta0, as.integer(niter), as.numeric(tol), as.integer(n.devices), PACKAGE = 'ldsr')
}

.pick <- function(res, g, series, jobs) {
  b <- res$best[g]
  if (is.na(b)) stop('no restart produced a finite likelihood')
  s <- series[[jobs[[g]]$series]]
  first <- sum(sapply(jobs[seq_len(g - 1)], function(j) length(j$init)))
  list(theta = .theta_list(res$theta[, b], nrow(s$u), nrow(s$v)),
       fit = list(X = res$X[[g]], Y = res$Y[[g]], V = res$V[[g]], J = res$J[[g]], lik = res$lik[b]),
       lik = res$lik[b], init = jobs[[g]]$init[[b - first]])
}

# R/LDS_reconstruction.R:42-62
LDS_EM_restart <- function(y, u, v, init, niter = 1000, tol = 1e-5, return.init = TRUE) {
  series <- list(list(y = y, u = u, v = v))
  jobs <- list(list(series = 1L, held = integer(0), init = init))
  res <- .em_batch(series, jobs, niter, tol, 1L)
  if (any(res$status == 1L)) stop('inv(): matrix is singular')
  ans <- .pick(res, 1L, series, jobs)
  if (!return.init) ans$init <- NULL
  ans
}

# R/LDS_reconstruction.R:270-285 (kept for callers that use it directly)
one_lds_cv <- function(z, instPeriod, mu, y, u, v, method = 'EM', num.restarts = 20,
                       ub = NULL, lb = NULL, num.islands = 4, pop.per.island = 100,
                       niter = 1000, tol = 1e-6, use.raw = FALSE) {
  stopifnot(method == 'EM')
  y[instPeriod][z] <- NA
  result <- LDS_EM_restart(y, u, v, make_init(nrow(u), nrow(v), num.restarts), niter, tol, FALSE)
  if (use.raw) c(propagate(result$theta, u, v, y)$Y[instPeriod]) + mu
  else c(result$fit$Y[instPeriod]) + mu
}

# The fold loop of cvLDS (R/LDS_reconstruction.R:372-382) as one batch.  Call it in place of the
# `Ycv <- if (single) foreach(...) else foreach(...) %:% foreach(...)` block; everything before
# (:308-370) and after (:384-409) stays as it is.
cv_folds_batched <- function(Z, instPeriod, mu, y, u, v, num.restarts, niter, tol, n.devices = 0L) {
  single <- is.matrix(u)
  if (single) { u <- list(u); v <- list(v) }
  series <- lapply(seq_along(u), function(i) list(y = y, u = u[[i]], v = v[[i]]))
  jobs <- list()
  for (z in Z) for (i in seq_along(u))   # fold-major, member-minor: the reference's nested foreach order
    jobs[[length(jobs) + 1L]] <- list(series = i, held = instPeriod[z],
                                      init = make_init(nrow(u[[i]]), nrow(v[[i]]), num.restarts))
  res <- .em_batch(series, jobs, niter, tol, n.devices)
  if (any(res$status == 1L)) stop('inv(): matrix is singular')
  nm <- length(u)
  lapply(seq_along(Z), function(k) {
    cols <- sapply(seq_len(nm), function(i) {
      g <- (k - 1L) * nm + i
      if (is.na(res$best[g])) stop('no restart produced a finite likelihood')
      c(res$Y[[g]][instPeriod]) + mu           # fit$Y[instPeriod] + mu   (:283)
    })
    rowMeans(matrix(cols, ncol = nm))           # .final = rowMeans         (:379)
  })
}

# The ensemble loop of LDS_reconstruction (R/LDS_reconstruction.R:236-246) as one batch: returns
# what `call_method(..., 'EM', ...)` returns for each member.
reconstruct_members_batched <- function(y, u, v, init, niter, tol, return.init, n.devices = 0L) {
  single <- is.matrix(u)
  if (single) { u <- list(u); v <- list(v); init <- list(init) }
  series <- lapply(seq_along(u), function(i) list(y = y, u = u[[i]], v = v[[i]]))
  jobs <- lapply(seq_along(u), function(i) list(series = i, held = integer(0), init = init[[i]]))
  res <- .em_batch(series, jobs, niter, tol, n.devices)
  if (any(res$status == 1L)) stop('inv(): matrix is singular')
  out <- lapply(seq_along(u), function(i) {
    a <- .pick(res, i, series, jobs)
    if (!return.init) a$init <- NULL
    a
  })
  if (single) out[[1]] else out
}

# R/stochastics.R:58-63.  exact.rng = TRUE (default): the noise comes from R's own RNG in exactly the
# order one_LDS_rep draws it -- per replicate rnorm(1) for x1, rnorm(n) for the state, rnorm(n) for the
# observations (R/stochastics.R:23-26); one rnorm() call of the total length yields the same stream as
# those calls in sequence, and rnorm(k, 0, s) is s * N(0,1) -- so set.seed() reproduces the
# reference's replicates.  r.seed = s: the same replicates as `set.seed(s); LDS_rep(...)` with R's
# default generators, but the stream (Mersenne-Twister + inversion) is generated on the device: no
# 1.3 GB of host noise for 100 000 replicates and no time in rnorm; R's own RNG state is left alone.
# exact.rng = FALSE uses the counter-based device generator keyed by `seed` (statistically, not
# bitwise, equal).
LDS_rep <- function(theta, u = NULL, v = NULL, years, num.reps = 100, mu = 0, exp.trans = TRUE,
                    exact.rng = TRUE, seed = 0, r.seed = NULL) {
  n <- length(years)
  z <- if (exact.rng && is.null(r.seed)) stats::rnorm(num.reps * (1 + 2 * n)) else NULL
  m <- .Call('_ldsr_rep_batch', theta, u, v, as.integer(n), as.integer(num.reps), as.numeric(seed),
             as.numeric(mu), as.logical(exp.trans), z, if (is.null(r.seed)) NULL else as.integer(r.seed),
             PACKAGE = 'ldsr')
  data.table::data.table(year = rep(years, num.reps), simX = m[, 1], simY = m[, 2], simQ = m[, 3],
                         rep = rep(seq_len(num.reps), each = n))
}


#' Kalman / RTS smoother for a state of dimension d > 1 (beyond ldsr, whose state is scalar).
#' theta: list(A d x d, B d x p, C 1 x d, D 1 x q, Q d x d, R, mu1 d x 1, V1 d x d); y 1 x T with NA.
#' method: 1 = associative scan over time (long series), 0 = sequential recursion.
#' Returns list(X d x T, Y 1 x T, V (d*d) x T with column t = vec(V_t), lik).
Kalman_smoother_d <- function(y, u, v, theta, stdlik = TRUE, method = 1L) {
  if (is.null(u)) u <- matrix(0)
  if (is.null(v)) v <- matrix(0)
  .Call(`_ldsr_smoother_d`, y, u, v, theta, stdlik, as.integer(method))
}


#' All folds' skill metrics in one device call; replaces
#' `mapply(calculate_metrics, sim = Ycv, z = Z, MoreArgs = list(obs = target))` in cvLDS
#' (R/LDS_reconstruction.R:395).  Ycv: list of per-fold vectors (or an n x n_folds matrix).
cv_metrics_batched <- function(Ycv, Z, target, exp_trans = FALSE) {
  if (is.list(Ycv)) Ycv <- do.call(cbind, Ycv)
  m <- .Call(`_ldsr_cv_metrics`, Ycv, as.numeric(target), lapply(Z, as.integer), exp_trans)
  rownames(m) <- c("R2", "RE", "CE", "nRMSE", "KGE")
  m
}


#' construct_rec for all ensemble members in one device call (R/LDS_reconstruction.R:190-212) and the
#' year-wise ensemble mean of X and Q (:247-248).  fits: list of fit lists (X, V, Y as 1 x T matrices),
#' thetas: the matching theta lists.  Returns list(members = list of data.tables, mean = data.table).
construct_rec_batched <- function(fits, thetas, mu, transform, years, lambda = 0) {
  X <- sapply(fits, function(f) c(f$X)); V <- sapply(fits, function(f) c(f$V)); Y <- sapply(fits, function(f) c(f$Y))
  C <- sapply(thetas, function(t) c(t$C)); R <- sapply(thetas, function(t) c(t$R))
  tr <- match(transform, c("none", "log", "boxcox")) - 1L
  r <- .Call(`_ldsr_construct_rec`, as.matrix(X), as.matrix(V), as.matrix(Y), as.numeric(C), as.numeric(R),
             as.numeric(mu), as.integer(tr), as.numeric(lambda))
  a <- array(r$rec, dim = c(length(years), 6L, length(fits)))
  members <- lapply(seq_along(fits), function(i) {
    dt <- data.table::as.data.table(a[, , i]); data.table::setnames(dt, c("X", "Xl", "Xu", "Q", "Ql", "Qu"))
    cbind(data.table::data.table(year = years), dt)
  })
  list(members = members, mean = data.table::data.table(year = years, X = r$mean[, 1], Q = r$mean[, 2]))
}


#' Objective of the experimental learners for a whole population (R/LDS_GA.R:28-44, 136-147):
#' `thetas` is a (2d+6) x n matrix of candidate vectors; kind is "penalized_likelihood", "negLogLik" or
#' "ssqTrain".  GA::gaisl / optim call the scalar versions once per candidate; with a vectorised fitness
#' (e.g. GA's `parallel` hook or a custom generation loop) one generation is one device call.
objective_batched <- function(y, u, v, thetas, kind = "penalized_likelihood", lambda = 1) {
  k <- match(kind, c("penalized_likelihood", "negLogLik", "ssqTrain")) - 1L
  .Call(`_ldsr_objective`, y, u, v, thetas, as.integer(k), as.numeric(lambda))
}


#' The shim keeps one device context per R session; its device and pinned buffers are cached between
#' calls.  This hands the cached device memory back to the driver (bytes released, invisibly); the next
#' call simply allocates again.  Unloading the package (R_unload_ldsr) releases everything.
ldsr_gpu_trim <- function() invisible(.Call(`_ldsr_trim`))
