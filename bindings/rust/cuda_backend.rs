// Safe wrapper over cuda_ffi.rs for `crates/gp` (feature `cuda`): the handle type that `GaussianProcess` keeps next to
// its `GpInnerParams`, and the three calls `algorithm.rs` makes -- the objective (:880-897), the final evaluation
// (:966-968) and `predict*` (:253-307).  Hand-written companion of the generated declarations; it cannot be compiled in
// the build image (no Rust toolchain) and is the patch INTEGRATION.md describes, kept next to the library it binds.
#![cfg(feature = "cuda")]
use crate::cuda_ffi::*;
use crate::errors::{GpError, Result};
use ndarray::{Array1, Array2, ArrayBase, Data, Ix1, Ix2};
use std::ffi::CStr;
use std::os::raw::c_int;

pub struct CudaCtx {
    raw: *mut EgxGpCtx,
    n: usize,
    p: usize,
    slots: c_int,
}
// the library serialises calls per handle and gives every slot of the asynchronous seam its own workspace
unsafe impl Send for CudaCtx {}
unsafe impl Sync for CudaCtx {}

fn status(st: c_int) -> Result<()> {
    let msg = || unsafe { CStr::from_ptr(egx_last_error()).to_string_lossy().into_owned() };
    match st {
        EGX_OK => Ok(()),
        EGX_NOT_POSITIVE_DEFINITE => Err(GpError::LinalgError(linfa_linalg::LinalgError::NotPositiveDefinite)),
        EGX_ILL_CONDITIONED_FT => Err(GpError::LikelihoodComputationError(
            "ft is too ill conditioned, try another theta again".to_string(),
        )),
        EGX_ILL_CONDITIONED_F => Err(GpError::LikelihoodComputationError(
            "F is too ill conditioned. Poor combination of regression model and observations.".to_string(),
        )),
        EGX_INVALID_VALUE => Err(GpError::InvalidValueError(msg())),
        _ => Err(GpError::InvalidValueError(format!("CUDA error: {}", msg()))),
    }
}

impl CudaCtx {
    /// xnorm / ynorm: the normalised training set of `fit` (algorithm.rs:840-841); corr / mean: the correlation / mean model constants of cuda_ffi.rs.
    #[allow(clippy::too_many_arguments)]
    pub fn new(
        xnorm: &Array2<f64>, x_mean: &Array1<f64>, x_std: &Array1<f64>, ynorm: &Array1<f64>, y_mean: f64, y_std: f64,
        w_star: &Array2<f64>, corr: c_int, mean: c_int, nugget: f64, n_chains: usize,
    ) -> Result<Self> {
        let (n, d) = xnorm.dim();
        let mut raw = std::ptr::null_mut();
        let x = xnorm.as_standard_layout();
        let w = w_star.as_standard_layout();
        status(unsafe {
            egx_gp_create(&mut raw, 0, corr, mean, x.as_ptr(), n as c_int, d as c_int, ynorm.as_ptr(), x_mean.as_ptr(),
                          x_std.as_ptr(), y_mean, y_std, w.as_ptr(), w_star.ncols() as c_int, nugget)
        })?;
        let mut p = 0;
        unsafe { egx_gp_dims(raw, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), &mut p) };
        let slots = unsafe { egx_gp_async_slots(raw, n_chains as c_int) };
        Ok(CudaCtx { raw, n, p: p as usize, slots })
    }

    /// The body of `objfn` (algorithm.rs:880-897): -rlf, +inf on any error.  `worker` = rayon::current_thread_index().
    pub fn objective(&self, theta: &[f64], worker: usize) -> f64 {
        let mut rlf = f64::NAN;
        let st = if self.slots >= 2 {
            let slot = (worker % self.slots as usize) as c_int;      // callers hold a per-slot lock
            unsafe {
                match egx_gp_eval_begin(self.raw, slot, theta.as_ptr()) {
                    EGX_OK => egx_gp_eval_end(self.raw, slot, &mut rlf),
                    e => e,
                }
            }
        } else {
            unsafe { egx_gp_reduced_likelihood(self.raw, theta.as_ptr(), &mut rlf) }
        };
        if st == EGX_OK && !rlf.is_nan() { -rlf } else { f64::INFINITY }
    }

    /// The final evaluation (algorithm.rs:966-968): likelihood and the pieces of `GpInnerParams` (r_chol on demand).
    pub fn finalize(&self, theta: &[f64]) -> Result<(f64, f64, Array1<f64>, Array1<f64>, Array2<f64>, Array2<f64>)> {
        let (mut rlf, mut sigma2) = (f64::NAN, f64::NAN);
        let mut beta = Array1::<f64>::zeros(self.p);
        let mut gamma = Array1::<f64>::zeros(self.n);
        let mut ft = Array2::<f64>::zeros((self.n, self.p));
        let mut g = Array2::<f64>::zeros((self.p, self.p));
        status(unsafe {
            egx_gp_finalize(self.raw, theta.as_ptr(), &mut rlf, &mut sigma2, beta.as_mut_ptr(), gamma.as_mut_ptr(),
                            ft.as_mut_ptr(), g.as_mut_ptr())
        })?;
        Ok((rlf, sigma2, beta, gamma, ft, g))
    }

    pub fn r_chol(&self) -> Result<Array2<f64>> {
        let mut l = Array2::<f64>::zeros((self.n, self.n));
        status(unsafe { egx_gp_download_chol(self.raw, l.as_mut_ptr()) })?;
        Ok(l)
    }

    /// predict_valvar (algorithm.rs:282-307) on raw inputs.
    pub fn predict_valvar(&self, x: &ArrayBase<impl Data<Elem = f64>, Ix2>) -> Result<(Array1<f64>, Array1<f64>)> {
        let xs = x.as_standard_layout();
        let m = xs.nrows();
        let (mut y, mut v) = (Array1::<f64>::zeros(m), Array1::<f64>::zeros(m));
        status(unsafe { egx_gp_predict_valvar(self.raw, xs.as_ptr(), m as c_int, y.as_mut_ptr(), v.as_mut_ptr()) })?;
        Ok((y, v))
    }

    pub fn predict(&self, x: &ArrayBase<impl Data<Elem = f64>, Ix2>) -> Result<Array1<f64>> {
        let xs = x.as_standard_layout();
        let mut y = Array1::<f64>::zeros(xs.nrows());
        status(unsafe { egx_gp_predict(self.raw, xs.as_ptr(), xs.nrows() as c_int, y.as_mut_ptr()) })?;
        Ok(y)
    }

    pub fn predict_var(&self, x: &ArrayBase<impl Data<Elem = f64>, Ix2>) -> Result<Array1<f64>> {
        let xs = x.as_standard_layout();
        let mut v = Array1::<f64>::zeros(xs.nrows());
        status(unsafe { egx_gp_predict_var(self.raw, xs.as_ptr(), xs.nrows() as c_int, v.as_mut_ptr()) })?;
        Ok(v)
    }
}

impl Drop for CudaCtx {
    fn drop(&mut self) {
        unsafe { egx_gp_destroy(self.raw) }
    }
}

#[allow(dead_code)]
fn _assert_bounds(_: &ArrayBase<impl Data<Elem = f64>, Ix1>) {}
