/* The scenario of the reference's crates/gp/examples/kriging.rs (5 training points, default Kriging = constant mean +
 * squared exponential, 100 predictions on [0, 4]) in plain C99 through the C ABI -- what a cgo / JNI / ctypes / Rust FFI
 * caller does.
 *   gcc -std=c99 -Iinclude examples/kriging.c -Legobox_b200 -legobox_gpu -Wl,-rpath,$PWD/egobox_b200 -lm -o kriging
 * Exit code 0 = fitted and predicted on the GPU; 2 = the library reported an error (e.g. no CUDA device). */
#include <stdio.h>
#include <stdlib.h>

#include "egobox_gpu.h"

int main(void) {
    const double xtrain[5] = {0.0, 1.0, 2.0, 3.0, 4.0};
    const double ytrain[5] = {0.0, 1.0, 1.5, 0.9, 1.0};
    egx_gp_params prm;
    egx_gp_model* gp = NULL;
    double xtest[100], ytest[100], vtest[100], theta[1];
    int i, st;

    egx_gp_params_default(&prm);                 /* Kriging::params(): GpValidParams::default, parameters.rs:105-120 */
    st = egx_gp_fit(&prm, xtrain, 5, 1, ytrain, &gp);
    if (st != EGX_OK) {
        fprintf(stderr, "Kriging fitting: status %d: %s\n", st, egx_last_error());
        return 2;
    }
    for (i = 0; i < 100; ++i) xtest[i] = 4.0 * i / 99.0;
    st = egx_gp_model_predict_valvar(gp, xtest, 100, ytest, vtest);
    if (st != EGX_OK) {
        fprintf(stderr, "Kriging prediction: status %d: %s\n", st, egx_last_error());
        egx_gp_model_destroy(gp);
        return 2;
    }
    egx_gp_model_theta(gp, theta);
    printf("theta %.6f likelihood %.9f variance %.6f evals %lld\n", theta[0], egx_gp_model_likelihood(gp),
           egx_gp_model_variance(gp), egx_gp_model_n_evals(gp));
    for (i = 0; i < 100; i += 33) printf("x %.4f  y %.6f  var %.3e\n", xtest[i], ytest[i], vtest[i]);
    egx_gp_model_destroy(gp);
    return 0;
}
