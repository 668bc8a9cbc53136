# R-side glue for the GPU shim (replaces R/RcppExports.R of statespacer; same five stubs, same arguments --
# R/RcppExports.R:11-99 -- plus the batched objective).  Not run in this repository's image (no R there).

# The Rcpp signatures of the reference (NumericMatrix, arma::cube, LogicalMatrix: src/RcppExports.cpp:17-105) coerce
# integer-typed arguments -- statespacer() never coerces y, and count data arrive as an integer matrix.  The C shim reads
# REAL() / LOGICAL() directly, so the stubs do the coercion (dim and dimnames are kept by storage.mode<-).
.d <- function(x) { if (!is.null(x) && !is.double(x)) storage.mode(x) <- "double"; x }
.l <- function(x) { if (!is.logical(x)) storage.mode(x) <- "logical"; x }

LogLikC <- function(y, y_isna, a, P_inf, P_star, Z, T, R, Q) {
  .Call(`_statespacer_LogLikC`, .d(y), .l(y_isna), .d(a), .d(P_inf), .d(P_star), .d(Z), .d(T), .d(R), .d(Q))
}
KalmanC <- function(y, y_isna, a, P_inf, P_star, Z, T, R, Q, diagnostics) {
  .Call(`_statespacer_KalmanC`, .d(y), .l(y_isna), .d(a), .d(P_inf), .d(P_star), .d(Z), .d(T), .d(R), .d(Q),
        as.logical(diagnostics))
}
FastSmootherC <- function(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, transposed_state) {
  .Call(`_statespacer_FastSmootherC`, .d(y), .d(a), .d(P_inf), .d(P_star), .d(Z), .d(T), .d(R), .d(Q),
        as.integer(initialisation_steps), as.logical(transposed_state))
}
SimulateC <- function(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only, transposed_state) {
  .Call(`_statespacer_SimulateC`, as.integer(nsim), as.integer(repeat_Q), as.integer(N), .d(a), .d(Z), .d(T), .d(R),
        .d(Q), .d(P_star), as.logical(draw_initial), as.logical(eta_only), as.logical(transposed_state))
}
ExtractComponentC <- function(a, Z) {
  .Call(`_statespacer_ExtractComponentC`, .d(a), .d(Z))
}

#' Batched loglikelihood: B series x V parameter vectors in one launch.
#'
#' @param Y B x N matrix (one series per row, p = 1) or B x (N*p) with the p observations of a time point
#'   adjacent; NA = missing.  Stored column-major by R, i.e. already time-major and batch-contiguous.
#' @param V number of parameter vectors per series (1 + 2k for a central-difference gradient).
#' @param Q_diag,P_star_diag r x (B*V) and m x (B*V) matrices of per-evaluation diagonals
#'   (column e = (v-1)*B + b), or NULL for the shared Q / P_star of the system matrices.
#' @return B x V matrix of loglik / N (NA where LogLikC would return NA).
LogLikBatchC <- function(Y, V, a, P_inf, P_star, Z, T, R, Q, Q_diag = NULL, P_star_diag = NULL) {
  Y <- .d(Y)
  Y[is.na(Y)] <- NaN
  .Call(`_statespacer_LogLikBatchC`, Y, as.integer(V), .d(a), .d(P_inf), .d(P_star), .d(Z), .d(T), .d(R), .d(Q),
        .d(Q_diag), .d(P_star_diag))
}

#' Objective + central-difference gradient for optim(gr = ...) from ONE launch (replaces the 2k sequential
#' objective calls stats::optim makes per gradient, R/statespacer.R:977-982).  `sysmat_diag(theta)` returns the
#' Q and P_star diagonals of the model at theta (e.g. exp(2 * theta) placed per R/statespacer.R:444-509).
loglik_and_gradient <- function(theta, y, sysmat, sysmat_diag, h = 1e-3) {
  k <- length(theta)
  thetas <- cbind(theta, do.call(cbind, lapply(seq_len(k), function(i) {
    e <- replace(numeric(k), i, h); cbind(theta + e, theta - e) })))
  d <- lapply(seq_len(ncol(thetas)), function(j) sysmat_diag(thetas[, j]))
  # one row per series with the p observations of a time point ADJACENT (index t * p + j): y is N x p, stored
  # column-major (t + N * j), so the row is t(y) flattened -- matrix(y, nrow = 1) would scramble a multivariate y
  ll <- LogLikBatchC(matrix(t(as.matrix(y)), nrow = 1), ncol(thetas), sysmat$a, sysmat$P_inf, sysmat$P_star, sysmat$Z, sysmat$T,
                     sysmat$R, sysmat$Q, Q_diag = sapply(d, `[[`, "Q"), P_star_diag = sapply(d, `[[`, "P_star"))
  list(value = ll[1, 1], gradient = (ll[1, seq(2, 2 * k, 2)] - ll[1, seq(3, 2 * k + 1, 2)]) / (2 * h))
}

#' Objective and gradient straight from parameter vectors: the variances exp(2 * theta) are placed on the
#' diagonals of Q / P_star on the GPU (q_index / p_star_index: 1-based parameter index per diagonal entry, 0 = keep
#' the entry of the shared matrix), so R neither rebuilds system matrices nor ships them per evaluation.
#' With one series this is optim(fn =, gr =) for statespacer()'s loop (R/statespacer.R:917-982); with B series it
#' fits B models at once.
#' @param Y B x (N*p) matrix as for LogLikBatchC; @param theta k x B matrix.
#' @return list(value = numeric(B), gradient = k x B matrix), both of loglik / N.
LogLikGradBatchC <- function(Y, theta, q_index, p_star_index, sysmat, h = 1e-3) {
  Y <- .d(Y)
  Y[is.na(Y)] <- NaN
  .Call(`_statespacer_LogLikGradBatchC`, Y, .d(as.matrix(theta)), as.integer(q_index) - 1L,
        as.integer(p_star_index) - 1L, as.numeric(h), .d(sysmat$a), .d(sysmat$P_inf), .d(sysmat$P_star), .d(sysmat$Z),
        .d(sysmat$T), .d(sysmat$R), .d(sysmat$Q))
}

#' SimSmoother() as one device-resident call (replaces the body of R/SimSmoother.R:53-766).
#'
#' `comps` lists, in the order R/SimSmoother.R calls SimulateC, the arguments of each of those calls
#' (a, Z, T, R, Q, P_star, repeat_Q, draw_initial; 3-D arrays with a trailing dim); `full` is the model of
#' R/SimSmoother.R:579-615 (residual states in front); `Z_extract` the Z_padded arrays of :645-766.  The normal variates
#' are drawn inside the shim with norm_rand() in the reference's order, so set.seed() reproduces the CPU package's draws.
#' @return list(epsilon = N x p x nsim, a = N x m x nsim, eta = N x r x nsim, components = list of N x p x nsim).
SimSmootherC <- function(nsim, N, p, comps, H, full, initialisation_steps, smoothed, Z_extract = list()) {
  if (is.matrix(H)) dim(H) <- c(dim(H), 1)
  .Call(`_statespacer_SimSmootherC`, as.integer(nsim), as.integer(N), as.integer(p), comps, H, full,
        as.integer(initialisation_steps), smoothed$a, smoothed$epsilon, smoothed$eta, Z_extract)
}

#' Collapse transform of R/statespacer.R:815-900 in one call (replaces the four branches there):
#' list(y = collapsed observations with NA, H_star, loglik_add).
CollapseC <- function(y, H, Z) .Call(`_statespacer_CollapseC`, y, H, Z)

#' Forecast recursion of R/predict.R:205-246: list(y_fc, a_fc, P_fc, Fmat_fc) from object$predicted$a_fc / P_fc.
ForecastC <- function(forecast_period, a_fc, P_fc, Z, T, R, Q, H) {
  .Call(`_statespacer_ForecastC`, as.integer(forecast_period), a_fc, P_fc, Z, T, R, Q, H)
}

#' predict()'s simulations over the forecast period (R/predict.R:142-856) as one call: the same shim entry with
#' full = NULL -- no smoothing, no mean correction; `comps` start from object$predicted$a_fc / P_fc with
#' draw_initial = TRUE.  Returns list(y, epsilon, a, eta, components).
SimulateModelC <- function(nsim, forecast_period, p, comps, H, Z_extract = list()) {
  if (is.matrix(H)) dim(H) <- c(dim(H), 1)
  .Call(`_statespacer_SimSmootherC`, as.integer(nsim), as.integer(forecast_period), as.integer(p), comps, H, NULL, 0L,
        NULL, NULL, NULL, Z_extract)
}

#' Use n GPUs (0 = all visible) for the batched objective from this R session: LogLikGradBatchC() then shards its series
#' over them inside the call (one worker thread and stream per device, no collective).  Returns the number in use.
UseDevices <- function(n = 0L) .Call(`_statespacer_UseDevices`, as.integer(n))
