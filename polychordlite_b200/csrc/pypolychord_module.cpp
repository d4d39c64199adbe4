// pypolychord_module.cpp -- the CPython extension `_pypolychord` of the B200 engine.
//
// Replaces /root/reference/pypolychord/_pypolychord.cpp:119-229 (+ _array.hpp, _pypolychord.hpp, _python.hpp): one
// module-level function `run` taking the reference's 34 positional arguments in the reference's order
// (format string _pypolychord.cpp:129, argument list polychord.py:600-634), three C trampolines that wrap engine
// buffers zero-copy as numpy arrays and call the Python callables (:29-112), and run_polychord(ll, prior, dumper,
// Settings) of the C++ facade underneath (:219).  Built in-tree against libchord.so (polychordlite_b200/_build.py);
// NOT part of libchord.so, exactly as the reference builds its shim as a separate extension (setup.py:114-123).
//
// Deliberate differences from the reference shim:
//   * the boolean settings are parsed into ints and then assigned (the reference lets PyArg_ParseTuple write 4-byte
//     ints into 1-byte bool members);
//   * a Python exception raised inside the prior or the dumper stops the run at once (the reference drops the NULL
//     return and the error surfaces later, :76, :111);
//   * a callable that carries a device form (an object with a `_pc_c_callback(nDims)` method returning the address of
//     a C callback registered through pc_register_device_likelihood / pc_register_device_prior, e.g.
//     pypolychord.builtin.Gaussian) is passed to the engine as that C pointer, so the whole sampling loop stays on
//     the GPU; any other callable takes the trampolines (the engine's host-callback path);
//   * the references taken on the callables are released on every path.
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <Python.h>
#include <numpy/arrayobject.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/interfaces.hpp"

namespace {

// Thrown by a trampoline when the Python error indicator is set: unwinds through the engine (which is
// exception-transparent for foreign types, csrc/pc_errors.h) back to run().
struct PythonErrorSet {};

// One run at a time per process, on the calling thread with the GIL held -- the reference's contract
// (_pypolychord.cpp:27,61,83 keep the callables in statics too).
struct Callables {
    PyObject* loglikelihood = nullptr;
    PyObject* prior = nullptr;
    PyObject* dumper = nullptr;
} g_py;

struct Ref {   // owned reference
    PyObject* o;
    explicit Ref(PyObject* p = nullptr) : o(p) {}
    Ref(const Ref&) = delete;
    Ref& operator=(const Ref&) = delete;
    ~Ref() { Py_XDECREF(o); }
    explicit operator bool() const { return o != nullptr; }
};

// zero-copy view of engine-owned memory (valid during the callback only)
PyObject* view(double* data, int nd, npy_intp d0, npy_intp d1, bool writeable) {
    npy_intp shape[2] = {d0, d1};
    static double nothing = 0.0;   // a NULL data pointer makes numpy allocate: give empty arrays an address
    PyObject* a = PyArray_SimpleNewFromData(nd, shape, NPY_DOUBLE, data ? (void*)data : (void*)&nothing);
    if (!a) throw PythonErrorSet();
    if (!writeable) PyArray_CLEARFLAGS(reinterpret_cast<PyArrayObject*>(a), NPY_ARRAY_WRITEABLE);
    return a;
}

// loglikelihood(theta, phi) -> float; phi is filled in place (_pypolychord.cpp:29-58)
double ll_trampoline(double* theta, int nDims, double* phi, int nDerived) {
    Ref a_theta(view(theta, 1, nDims, 0, false));
    Ref a_phi(view(phi, 1, nDerived, 0, true));
    Ref res(PyObject_CallFunctionObjArgs(g_py.loglikelihood, a_theta.o, a_phi.o, nullptr));
    if (!res) throw PythonErrorSet();
    if (!PyFloat_Check(res.o)) {
        PyErr_SetString(PyExc_TypeError, "loglikelihood must be a float (element 0 of loglikelihood return)");
        throw PythonErrorSet();
    }
    return PyFloat_AsDouble(res.o);
}

// prior(cube, theta): theta is filled in place (_pypolychord.cpp:63-80)
void prior_trampoline(double* cube, double* theta, int nDims) {
    Ref a_cube(view(cube, 1, nDims, 0, false));
    Ref a_theta(view(theta, 1, nDims, 0, true));
    Ref res(PyObject_CallFunctionObjArgs(g_py.prior, a_cube.o, a_theta.o, nullptr));
    if (!res) throw PythonErrorSet();
}

// dumper(live, dead, logweights, logZ, logZerr); arrays are C row-major (npoints, npars) (_pypolychord.cpp:85-112)
void dumper_trampoline(int ndead, int nlive, int npars, double* live, double* dead, double* logweights, double logZ,
                       double logZerr) {
    Ref a_live(view(live, 2, nlive, npars, false));
    Ref a_dead(view(dead, 2, ndead, npars, false));
    Ref a_lw(view(logweights, 1, ndead, 0, false));
    Ref z(PyFloat_FromDouble(logZ)), ze(PyFloat_FromDouble(logZerr));
    if (!z || !ze) throw PythonErrorSet();
    Ref res(PyObject_CallFunctionObjArgs(g_py.dumper, a_live.o, a_dead.o, a_lw.o, z.o, ze.o, nullptr));
    if (!res) throw PythonErrorSet();
}

// obj._pc_c_callback(nDims) -> address of a C callback with a device form, or 0 when the object has none
bool device_pointer(PyObject* obj, int nDims, std::uintptr_t& out) {
    out = 0;
    if (!PyObject_HasAttrString(obj, "_pc_c_callback")) return true;
    Ref r(PyObject_CallMethod(obj, "_pc_c_callback", "i", nDims));
    if (!r) return false;
    if (r.o == Py_None) return true;
    out = (std::uintptr_t)PyLong_AsUnsignedLongLong(r.o);
    return !PyErr_Occurred();
}

bool list_to_doubles(PyObject* list, std::vector<double>& out) {   // _array.hpp:7-19
    const Py_ssize_t n = PyList_Size(list);
    out.clear();
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* it = PyList_GET_ITEM(list, i);
        if (!it || !PyNumber_Check(it)) return false;
        const double v = PyFloat_AsDouble(it);
        if (v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); return false; }
        out.push_back(v);
    }
    return true;
}
bool list_to_ints(PyObject* list, std::vector<int>& out) {   // _array.hpp:20-38
    const Py_ssize_t n = PyList_Size(list);
    out.clear();
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* it = PyList_GET_ITEM(list, i);
        if (!it || !PyLong_Check(it)) return false;
        out.push_back((int)PyLong_AsLong(it));
    }
    return true;
}
bool dict_to_schedule(PyObject* dict, std::vector<double>& loglikes, std::vector<int>& nlives) {   // _array.hpp:40-63
    PyObject *key, *value;
    Py_ssize_t pos = 0;
    loglikes.clear();
    nlives.clear();
    while (PyDict_Next(dict, &pos, &key, &value)) {
        if (!PyFloat_Check(key) || !PyLong_Check(value)) return false;
        loglikes.push_back(PyFloat_AsDouble(key));
        nlives.push_back((int)PyLong_AsLong(value));
    }
    return true;
}

PyObject* run(PyObject*, PyObject* args) {
    Settings S;
    PyObject *ll = nullptr, *prior = nullptr, *dumper = nullptr, *grade_frac = nullptr, *grade_dims = nullptr, *nlives = nullptr;
    const char *base_dir = nullptr, *file_root = nullptr;
    // the booleans of polychord.py:600-634, parsed as ints
    int do_clustering, posteriors, equals, cluster_posteriors, write_resume, write_paramnames, read_resume, write_stats,
        write_live, write_dead, write_prior, maximise, synchronous;
    if (!PyArg_ParseTuple(args, "OOOiiiiiiiiddidiiiiiiiiiiidissO!O!O!i:run", &ll, &prior, &dumper, &S.nDims, &S.nDerived,
                          &S.nlive, &S.num_repeats, &S.nprior, &S.nfail, &do_clustering, &S.feedback,
                          &S.precision_criterion, &S.logzero, &S.max_ndead, &S.boost_posterior, &posteriors, &equals,
                          &cluster_posteriors, &write_resume, &write_paramnames, &read_resume, &write_stats, &write_live,
                          &write_dead, &write_prior, &maximise, &S.compression_factor, &synchronous, &base_dir, &file_root,
                          &PyList_Type, &grade_frac, &PyList_Type, &grade_dims, &PyDict_Type, &nlives, &S.seed))
        return nullptr;
    S.do_clustering = do_clustering != 0; S.posteriors = posteriors != 0; S.equals = equals != 0;
    S.cluster_posteriors = cluster_posteriors != 0; S.write_resume = write_resume != 0;
    S.write_paramnames = write_paramnames != 0; S.read_resume = read_resume != 0; S.write_stats = write_stats != 0;
    S.write_live = write_live != 0; S.write_dead = write_dead != 0; S.write_prior = write_prior != 0;
    S.maximise = maximise != 0; S.synchronous = synchronous != 0;
    S.base_dir = base_dir;
    S.file_root = file_root;
    // the reference's argument checks and messages (_pypolychord.cpp:173-204)
    if (!list_to_doubles(grade_frac, S.grade_frac)) {
        PyErr_SetString(PyExc_TypeError, "grade_frac must be a list of doubles");
        return nullptr;
    }
    if (!list_to_ints(grade_dims, S.grade_dims)) {
        PyErr_SetString(PyExc_TypeError, "grade_dims must be a list of integers");
        return nullptr;
    }
    if (S.grade_frac.size() != S.grade_dims.size()) {
        PyErr_SetString(PyExc_ValueError, "grade_dims and grade_frac must have the same size");
        return nullptr;
    }
    long tot = 0;
    for (int d : S.grade_dims) tot += d;
    if (tot != S.nDims) {
        PyErr_SetString(PyExc_ValueError, "grade_dims must sum to nDims");
        return nullptr;
    }
    if (!dict_to_schedule(nlives, S.loglikes, S.nlives)) {
        PyErr_SetString(PyExc_TypeError, "nlives must be a dict mapping floats to integers");
        return nullptr;
    }
    if (!PyCallable_Check(ll) || !PyCallable_Check(prior) || (dumper != Py_None && !PyCallable_Check(dumper))) {
        PyErr_SetString(PyExc_TypeError, "loglikelihood, prior and dumper must be callable");
        return nullptr;
    }
    std::uintptr_t c_ll = 0, c_prior = 0;
    if (!device_pointer(ll, S.nDims, c_ll) || !device_pointer(prior, S.nDims, c_prior)) return nullptr;

    struct Hold {   // the callables stay alive for the run and are released on every path
        Hold(PyObject* a, PyObject* b, PyObject* c) {
            Py_INCREF(a); Py_INCREF(b); Py_INCREF(c);
            g_py.loglikelihood = a; g_py.prior = b; g_py.dumper = c;
        }
        ~Hold() {
            Py_XDECREF(g_py.loglikelihood); Py_XDECREF(g_py.prior); Py_XDECREF(g_py.dumper);
            g_py = Callables();
        }
    } hold(ll, prior, dumper);

    try {
        run_polychord(c_ll ? reinterpret_cast<pc_cxx_loglikelihood>(c_ll) : ll_trampoline,
                      c_prior ? reinterpret_cast<pc_cxx_prior>(c_prior) : prior_trampoline,
                      dumper == Py_None ? default_dumper : dumper_trampoline, S);
    } catch (const PythonErrorSet&) {
        return nullptr;   // the Python exception is already set
    }
    if (PyErr_Occurred()) return nullptr;
    Py_RETURN_NONE;
}

PyMethodDef methods[] = {{"run", run, METH_VARARGS, "Runs pypolychord on the B200 engine"}, {nullptr, nullptr, 0, nullptr}};
PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pypolychord",
                         "pypolychord: Python interface to the B200-native PolyChord engine (libchord.so).", -1, methods,
                         nullptr, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit__pypolychord(void) {
    import_array();
    return PyModule_Create(&moduledef);
}
