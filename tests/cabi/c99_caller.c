/* Plain C99 caller of the C ABI (tests/test_cabi_symbols.py::test_header_is_plain_c_and_links_from_c): the drop-in boundary must be
   usable without C++ or Python -- no torch types, no exceptions, status codes + pinn_last_error(). */
#include "pinn_b200.h"
#include <stdio.h>
int main(void) {
  pinn_t* h = 0;
  int layers[3] = {2, 20, 1};
  double lb[2] = {-1.0, 0.0}, ub[2] = {1.0, 1.0};
  printf("%s\n", pinn_version());
  int rc = pinn_create(&h, PINN_BURGERS_INF, 3, layers, lb, ub, 0, 0, 1, 0);
  printf("rc=%d err=%s\n", rc, pinn_last_error());
  return rc == 0 ? 1 : 0;   /* on a CPU-only box creation must fail loudly */
}
