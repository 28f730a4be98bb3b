/* Plain-C consumer of include/ode_b200.h: proves the header is C99-clean, the library links,
 * and prints struct sizes / defaults for the Python test to compare with its ctypes mirror. */
#include <stdio.h>
#include <string.h>
#include "ode_b200.h"

int main(void) {
    ode_b200_config c;
    if (ode_b200_config_new(ODE_B200_DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, &c) != 0) return 2;
    printf("abi %d sizeof_config %zu sizeof_outputs %zu devices %d\n", ode_b200_abi_version(), sizeof(ode_b200_config),
           sizeof(ode_b200_outputs), ode_b200_device_count());
    printf("alpha %.17g beta %.17g n_max %u lorenz %d dim %d\n", c.alpha, c.beta, c.n_max,
           ode_b200_system_builtin("lorenz"), ode_b200_system_dim(ode_b200_system_builtin("kepler")));
    if (ode_b200_device_count() == 0) {
        ode_b200_ctx* ctx = ode_b200_create(0);
        printf("create without gpu: %s (%s)\n", ctx ? "non-null" : "NULL", ode_b200_last_error());
        if (ctx) return 3;
    }
    return 0;
}
