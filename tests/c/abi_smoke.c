/* Plain-C consumer of include/ode_b200.h: proves the header is C99-clean, the library links,
 * and prints struct sizes / defaults for the Python test to compare with its ctypes mirror. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "ode_b200.h"

#define OFF(T, f) printf("offset " #T "." #f " %zu %zu\n", offsetof(T, f), sizeof(((T*)0)->f))

int main(void) {
    ode_b200_config c;
    if (ode_b200_config_new(ODE_B200_DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, &c) != 0) return 2;
    printf("abi %d sizeof_config %zu sizeof_outputs %zu devices %d\n", ode_b200_abi_version(), sizeof(ode_b200_config),
           sizeof(ode_b200_outputs), ode_b200_device_count());
    printf("alpha %.17g beta %.17g n_max %u lorenz %d dim %d\n", c.alpha, c.beta, c.n_max,
           ode_b200_system_builtin("lorenz"), ode_b200_system_dim(ode_b200_system_builtin("kepler")));
    /* every field of every ABI struct: offset and size, compared with the ctypes mirror (and with the
     * Rust #[repr(C)] mirror by tests/test_rust_ffi.py) */
    OFF(ode_b200_config, method); OFF(ode_b200_config, out_type); OFF(ode_b200_config, x0); OFF(ode_b200_config, x_end);
    OFF(ode_b200_config, dx); OFF(ode_b200_config, rtol); OFF(ode_b200_config, atol);
    OFF(ode_b200_config, safety_factor); OFF(ode_b200_config, alpha); OFF(ode_b200_config, beta);
    OFF(ode_b200_config, fac_min); OFF(ode_b200_config, fac_max); OFF(ode_b200_config, h_max); OFF(ode_b200_config, h0);
    OFF(ode_b200_config, uround); OFF(ode_b200_config, step_size); OFF(ode_b200_config, n_max);
    OFF(ode_b200_config, n_stiff);
    OFF(ode_b200_outputs, y_final); OFF(ode_b200_outputs, x_final); OFF(ode_b200_outputs, h_next);
    OFF(ode_b200_outputs, num_eval); OFF(ode_b200_outputs, accepted); OFF(ode_b200_outputs, rejected);
    OFF(ode_b200_outputs, n_step); OFF(ode_b200_outputs, status); OFF(ode_b200_outputs, out_cap);
    OFF(ode_b200_outputs, out_x); OFF(ode_b200_outputs, out_y); OFF(ode_b200_outputs, out_count);
    OFF(ode_b200_outputs, com_cap); OFF(ode_b200_outputs, com_lower); OFF(ode_b200_outputs, com_break);
    OFF(ode_b200_outputs, com_step); OFF(ode_b200_outputs, com_coef); OFF(ode_b200_outputs, com_count);
    OFF(ode_b200_outputs_f32, y_final); OFF(ode_b200_outputs_f32, x_final); OFF(ode_b200_outputs_f32, h_next);
    OFF(ode_b200_outputs_f32, num_eval); OFF(ode_b200_outputs_f32, accepted); OFF(ode_b200_outputs_f32, rejected);
    OFF(ode_b200_outputs_f32, n_step); OFF(ode_b200_outputs_f32, status); OFF(ode_b200_outputs_f32, out_cap);
    OFF(ode_b200_outputs_f32, out_x); OFF(ode_b200_outputs_f32, out_y); OFF(ode_b200_outputs_f32, out_count);
    OFF(ode_b200_outputs_f32, com_cap); OFF(ode_b200_outputs_f32, com_lower); OFF(ode_b200_outputs_f32, com_break);
    OFF(ode_b200_outputs_f32, com_step); OFF(ode_b200_outputs_f32, com_coef); OFF(ode_b200_outputs_f32, com_count);
    OFF(ode_b200_config_f32, method); OFF(ode_b200_config_f32, out_type); OFF(ode_b200_config_f32, x0);
    OFF(ode_b200_config_f32, x_end); OFF(ode_b200_config_f32, dx); OFF(ode_b200_config_f32, rtol);
    OFF(ode_b200_config_f32, atol); OFF(ode_b200_config_f32, safety_factor); OFF(ode_b200_config_f32, alpha);
    OFF(ode_b200_config_f32, beta); OFF(ode_b200_config_f32, fac_min); OFF(ode_b200_config_f32, fac_max);
    OFF(ode_b200_config_f32, h_max); OFF(ode_b200_config_f32, h0); OFF(ode_b200_config_f32, uround);
    OFF(ode_b200_config_f32, step_size); OFF(ode_b200_config_f32, n_max); OFF(ode_b200_config_f32, n_stiff);
    printf("sizeof_outputs_f32 %zu\n", sizeof(ode_b200_outputs_f32));
    if (ode_b200_device_count() == 0) {
        ode_b200_ctx* ctx = ode_b200_create(0);
        printf("create without gpu: %s (%s)\n", ctx ? "non-null" : "NULL", ode_b200_last_error());
        if (ctx) return 3;
    }
    return 0;
}
