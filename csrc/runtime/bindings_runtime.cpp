#include <pybind11/pybind11.h>
namespace py = pybind11;
namespace ssb {
void bind_runtime(py::module_& m) { (void)m; }
}  // namespace ssb
