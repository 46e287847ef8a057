// tcgen05 / TMEM / TMA matmul (modes 0 and 1 of tfcuda_matmul).  Placeholder until the kernel lands:
// it fails loudly instead of silently using another path.
#include "tfcuda_internal.h"

int tfcuda_matmul_tcgen05(const float*, const float*, float*, size_t, size_t, size_t, size_t, int mode) {
	tfcuda::set_error("tfcuda_matmul: tcgen05 path (mode " + std::to_string(mode) + ") is not built yet; use mode 2");
	return 1;
}
