// TEST INFRASTRUCTURE ONLY -- what utils_lib.cpp uses of pybind11: a string-keyed dict of numbers read through
// py::float_ / py::int_, and the module macro (its body is never instantiated here: the C shim calls
// generate_depth directly).
#pragma once
#include <map>
#include <string>

namespace pybind11 {
struct value { double v; };
struct float_ {
    double v;
    float_(const value& x) : v(x.v) {}
    operator float() const { return (float)v; }
};
struct int_ {
    long v;
    int_(const value& x) : v((long)x.v) {}
    operator int() const { return (int)v; }
};
struct dict {
    std::map<std::string, value> m;
    value& operator[](const char* k) { return m[k]; }
    bool contains(const char* k) const { return m.count(k) != 0; }
};
}  // namespace pybind11
#define PYBIND11_MODULE(name, var) template <class DpvUnusedModule> void dpv_unused_module_##name(DpvUnusedModule& var)
