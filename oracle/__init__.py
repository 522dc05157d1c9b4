"""CPU oracle for the particle -> grid scatter path.

TEST INFRASTRUCTURE ONLY (see sph_oracle.c).  Importable from tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs; the product
package ``sarracen_b200`` never imports it.
"""
from .oracle import (OracleBackend, OracleKernel, build, column_kernel, kernel_w,  # noqa: F401
                     line_int, num_threads, set_num_threads, surface_int)
