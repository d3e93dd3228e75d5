// ggpzero_interface — the reference's extension module (src/cpp/ggpzero_interface.cpp:50-96), for CPython 3,
// bound to libgz_b200's C-ABI (include/gz_b200.h). Same module name, same three factories, same argument
// formats, same method names:
//
//   GdlBasesTransformer(channel_size, channels_per_state, num_control_channels, num_prev_states, num_rewards,
//                       [policy_sizes])                    "iiiiiO!"   gdltransformer_impl.cpp:186-219
//       .add_board_base(base, offset) "ii"  .add_control_base(base, channel, value) "iif"  .test(sm_ptr) "n"
//   Supervisor(sm_ptr, transformer, batch_size, identifier)                "nO!is"     supervisor_impl.cpp:215-237
//       .start_self_play(num_workers, conf_dict) "iO!"   .poll(n, [arrays]) "iO!"   .fetch_samples()
//       .add_unique_state(bs_ptr) "n"   .clear_unique_states()
//   Player(sm_ptr, transformer, puct_conf_dict)                            "nO!O!"     player_impl.cpp:183-201
//       .player_reset "i"  .player_apply_move "n"  .player_move "nid"  .player_get_move "i"
//       .player_update_config "dii"  .player_balance_moves "i"  .player_tree_debug "i"  .poll "iO!"
//   buf_to_tuple_reverse_bytes(bytes)   GGPZeroInterfaceError
//
// Pointers cross as Python ints, as in the reference (cppinterface.py:12-30): sm_ptr is the gz_propnet* of the
// device program (DeviceStateMachine.sm_ptr), a joint move is a pointer to one int32 per role, a base state a
// pointer to its packed uint32 words (ggplib_api.JointMove.ptr / BaseState.ptr). Where the reference leaks or
// mis-counts references (gdltransformer_impl.cpp:35,47,179-182) this module does the right thing.
// Everything below the argument parsing is native: csrc/gz_host.cu (gz_supervisor_*, gz_player_*).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/gz_b200.h"

static PyObject* GzError = nullptr;

static PyObject* fail() {
    PyErr_SetString(GzError, gz_last_error());
    return nullptr;
}

// ---- configuration dicts -> the flat arrays of gz_puct_create (key lists: common.cpp:48-158)
static const char* PUCT_KEYS[22] = {
    "puct_constant", "puct_constant_root", "dirichlet_noise_pct", "noise_policy_squash_pct", "noise_policy_squash_prob",
    "choose", "random_scale", "temperature", "depth_temperature_start", "depth_temperature_increment",
    "depth_temperature_stop", "depth_temperature_max", "fpu_prior_discount", "fpu_prior_discount_root",
    "top_visits_best_guess_converge_ratio", "think_time", "converged_visits", "batch_size", "use_legals_count_draw",
    "backup_finalised", "lookup_transpositions", "evaluation_multiplier_to_convergence"};
static const char* SP_KEYS[12] = {
    "oscillate_sampling_pct", "temperature_for_policy", "evals_per_move", "resign0_score_probability", "resign0_pct",
    "resign1_score_probability", "resign1_pct", "run_to_end_pct", "run_to_end_evals", "run_to_end_early_score",
    "run_to_end_minimum_game_depth", "abort_max_length"};

static bool dict_number(PyObject* d, const char* key, double* out) {
    PyObject* v = PyDict_GetItemString(d, key);           // borrowed
    if (!v) { PyErr_Format(GzError, "configuration lacks the key '%s'", key); return false; }
    if (PyUnicode_Check(v)) {                             // choose: "choose_top_visits" | "choose_temperature"
        const char* s = PyUnicode_AsUTF8(v);
        *out = (s && std::string(s) == "choose_temperature") ? 1.0 : 0.0;
        return true;
    }
    const double x = PyFloat_AsDouble(v);
    if (x == -1.0 && PyErr_Occurred()) return false;
    *out = x;
    return true;
}
static bool puct_array(PyObject* d, double* out) {
    for (int i = 0; i < 22; i++) if (!dict_number(d, PUCT_KEYS[i], &out[i])) return false;
    return true;
}

// list of float32 C-contiguous numpy arrays -> pointers (common.cpp:186-196)
static bool prediction_pointers(PyObject* list, std::vector<const float*>& ptrs) {
    const Py_ssize_t n = PyList_Size(list);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject* o = PyList_GetItem(list, i);
        if (!PyArray_Check(o) || PyArray_TYPE((PyArrayObject*)o) != NPY_FLOAT || !PyArray_IS_C_CONTIGUOUS((PyArrayObject*)o)) {
            PyErr_SetString(GzError, "predictions must be C-contiguous float32 numpy arrays");
            return false;
        }
        ptrs.push_back((const float*)PyArray_DATA((PyArrayObject*)o));
    }
    return true;
}

static PyObject* planes_array(const float* data, Py_ssize_t count) {
    npy_intp dims[1] = {count};
    return PyArray_SimpleNewFromData(1, dims, NPY_FLOAT, (void*)data);     // a view of engine-owned pinned memory
}

// ------------------------------------------------------------------------------------------------ transformer
struct PyTransformer {
    PyObject_HEAD
    gz_transformer* t;
};
static void transformer_dealloc(PyTransformer* self) {
    if (self->t) gz_transformer_destroy(self->t);
    Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyObject* transformer_add_board_base(PyTransformer* self, PyObject* args) {
    int base = 0, off = 0;
    if (!PyArg_ParseTuple(args, "ii", &base, &off)) return nullptr;
    if (gz_transformer_add_board_base(self->t, base, off) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* transformer_add_control_base(PyTransformer* self, PyObject* args) {
    int base = 0, channel = 0;
    float value = 0;
    if (!PyArg_ParseTuple(args, "iif", &base, &channel, &value)) return nullptr;
    if (gz_transformer_add_control_base(self->t, base, channel, value) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* transformer_total_size(PyTransformer* self, PyObject*) { return PyLong_FromLong(gz_transformer_total_size(self->t)); }

// .test(sm_ptr): the reference's random playout + encode loop (gdltransformer_impl.cpp:50-124) on one game;
// returns the planes of the last position as a 1-D float32 array.
static PyObject* transformer_test(PyTransformer* self, PyObject* args) {
    Py_ssize_t sm_ptr = 0;
    if (!PyArg_ParseTuple(args, "n", &sm_ptr)) return nullptr;
    gz_propnet* pn = (gz_propnet*)sm_ptr;
    const int sw = (int)gz_propnet_info(pn, GZ_INFO_STATE_WORDS), lw = (int)gz_propnet_info(pn, GZ_INFO_LEGAL_WORDS);
    const int R = (int)gz_propnet_info(pn, GZ_INFO_N_ROLES), nb = (int)gz_propnet_info(pn, GZ_INFO_N_BASES);
    if (gz_transformer_finalize(self->t, nb, 0) != GZ_OK) return fail();
    const int total = gz_transformer_total_size(self->t);
    uint32_t *d_state = nullptr, *d_prev = nullptr, *d_legal = nullptr;
    uint8_t *d_term = nullptr, *d_goals = nullptr;
    float* d_out = nullptr;
    cudaMalloc(&d_state, sw * 4); cudaMalloc(&d_prev, sw * 4); cudaMalloc(&d_legal, lw * 4);
    cudaMalloc(&d_term, 1); cudaMalloc(&d_goals, R); cudaMalloc(&d_out, (size_t)total * 4);
    cudaMemset(d_state, 0, sw * 4); cudaMemset(d_term, 1, 1);
    int rc = GZ_OK;
    for (int step = 0; step < 200 && rc == GZ_OK; step++) {
        cudaMemcpy(d_prev, d_state, sw * 4, cudaMemcpyDeviceToDevice);
        rc = gz_sm_playout_step(pn, 0, 1, 0x7E57, (uint64_t)step, d_state, d_legal, d_term, d_goals, nullptr, nullptr,
                                GZ_GOALS_ALWAYS, nullptr);
        if (rc == GZ_OK) rc = gz_encode(self->t, 1, sw, d_state, step ? d_prev : nullptr, nullptr, step ? 1 : 0, d_out, GZ_DTYPE_F32, nullptr);
    }
    PyObject* res = nullptr;
    if (rc == GZ_OK) {
        npy_intp dims[1] = {total};
        res = PyArray_SimpleNew(1, dims, NPY_FLOAT);
        cudaMemcpy(PyArray_DATA((PyArrayObject*)res), d_out, (size_t)total * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_state); cudaFree(d_prev); cudaFree(d_legal); cudaFree(d_term); cudaFree(d_goals); cudaFree(d_out);
    return rc == GZ_OK ? res : fail();
}
static PyMethodDef transformer_methods[] = {
    {"add_board_base", (PyCFunction)transformer_add_board_base, METH_VARARGS, "add_board_base(base_index, buffer_offset)"},
    {"add_control_base", (PyCFunction)transformer_add_control_base, METH_VARARGS, "add_control_base(base_index, channel, value)"},
    {"total_size", (PyCFunction)transformer_total_size, METH_NOARGS, "floats per encoded position"},
    {"test", (PyCFunction)transformer_test, METH_VARARGS, "test(sm_ptr): random playout + encode; planes of the last position"},
    {nullptr, nullptr, 0, nullptr}};
static PyTypeObject PyTransformerType = {PyVarObject_HEAD_INIT(nullptr, 0)};

static PyObject* make_transformer(PyObject*, PyObject* args) {
    int cs = 0, cps = 0, ncc = 0, npv = 0, nrw = 0;
    PyObject* sizes = nullptr;
    if (!PyArg_ParseTuple(args, "iiiiiO!", &cs, &cps, &ncc, &npv, &nrw, &PyList_Type, &sizes)) return nullptr;
    std::vector<int> ps;
    for (Py_ssize_t i = 0; i < PyList_Size(sizes); i++) ps.push_back((int)PyLong_AsLong(PyList_GetItem(sizes, i)));
    if (PyErr_Occurred()) return nullptr;
    gz_transformer* t = nullptr;
    if (gz_transformer_create(cs, cps, ncc, npv, nrw, ps.data(), (int)ps.size(), &t) != GZ_OK) return fail();
    PyTransformer* o = PyObject_New(PyTransformer, &PyTransformerType);
    if (!o) { gz_transformer_destroy(t); return nullptr; }
    o->t = t;
    return (PyObject*)o;
}

// ------------------------------------------------------------------------------------------------ supervisor
struct PySupervisor {
    PyObject_HEAD
    gz_supervisor* s;
    gz_propnet* pn;
    PyObject* transformer;      // kept alive: the engine reads its tables
    std::string* identifier;
};
static void supervisor_dealloc(PySupervisor* self) {
    if (self->s) gz_supervisor_destroy(self->s);
    Py_XDECREF(self->transformer);
    delete self->identifier;
    Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyObject* supervisor_start_self_play(PySupervisor* self, PyObject* args) {
    int workers = 0;
    PyObject* conf = nullptr;
    if (!PyArg_ParseTuple(args, "iO!", &workers, &PyDict_Type, &conf)) return nullptr;
    double sp[12], pc[22], rc[22];
    for (int i = 0; i < 12; i++) if (!dict_number(conf, SP_KEYS[i], &sp[i])) return nullptr;
    PyObject* a = PyDict_GetItemString(conf, "puct_config");
    PyObject* b = PyDict_GetItemString(conf, "run_to_end_puct_config");
    if (!a || !b || !PyDict_Check(a) || !PyDict_Check(b)) { PyErr_SetString(GzError, "puct_config / run_to_end_puct_config dicts missing"); return nullptr; }
    if (!puct_array(a, pc) || !puct_array(b, rc)) return nullptr;
    if (gz_supervisor_start_self_play(self->s, workers, sp, pc, rc) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* supervisor_poll(PySupervisor* self, PyObject* args) {
    int n = 0;
    PyObject* list = nullptr;
    if (!PyArg_ParseTuple(args, "iO!", &n, &PyList_Type, &list)) return nullptr;
    std::vector<const float*> ptrs;
    if (n > 0 && !prediction_pointers(list, ptrs)) return nullptr;
    const float* planes = nullptr;
    const int rows = gz_supervisor_poll(self->s, n, ptrs.empty() ? nullptr : ptrs.data(), (int)ptrs.size(), &planes);
    if (rows < 0) return fail();
    if (rows == 0) Py_RETURN_NONE;
    return planes_array(planes, (Py_ssize_t)rows * gz_supervisor_info(self->s, 2));
}
// record layout: SR_* in csrc/gz_puct_core.h; dict keys: supervisor_impl.cpp:75-118
static PyObject* supervisor_fetch_samples(PySupervisor* self, PyObject*) {
    const int64_t srec = gz_supervisor_info(self->s, 3);
    if (srec <= 0) Py_RETURN_NONE;
    const int64_t cap = 1 << 17;
    std::vector<int32_t> buf((size_t)cap * srec), man(cap);
    const int64_t n = gz_supervisor_fetch_samples(self->s, buf.data(), cap, man.data());
    if (n < 0) return fail();
    if (n == 0) Py_RETURN_NONE;
    const int R = (int)gz_propnet_info(self->pn, GZ_INFO_N_ROLES), nb = (int)gz_propnet_info(self->pn, GZ_INFO_N_BASES);
    const int sw = (int)gz_propnet_info(self->pn, GZ_INFO_STATE_WORDS);
    const bool many = gz_supervisor_info(self->s, 0) > 1;
    PyObject* out = PyList_New(n);
    auto bits = [&](const int32_t* w) {
        PyObject* l = PyList_New(nb);
        for (int b = 0; b < nb; b++) PyList_SET_ITEM(l, b, PyLong_FromLong(((uint32_t)w[b >> 5] >> (b & 31)) & 1u));
        return l;
    };
    for (int64_t i = 0; i < n; i++) {
        const int32_t* r = buf.data() + i * srec;
        const float* f = (const float*)r;
        const int lead = r[9], nc = r[11];
        PyObject* d = PyDict_New();
        auto put = [&](const char* k, PyObject* v) { PyDict_SetItemString(d, k, v); Py_DECREF(v); };
        put("state", bits(r + 24));
        PyObject* prevs = PyList_New(0);
        if (r[10]) { PyObject* p = bits(r + 24 + sw); PyList_Append(prevs, p); Py_DECREF(p); }
        put("prev_states", prevs);
        PyObject* pols = PyList_New(R);
        for (int ri = 0; ri < R; ri++) {
            PyObject* pl;
            if (ri == lead) {
                pl = PyList_New(nc);
                for (int c = 0; c < nc; c++)
                    PyList_SET_ITEM(pl, c, Py_BuildValue("(if)", r[24 + 2 * sw + 2 * c], f[24 + 2 * sw + 2 * c + 1]));
            } else {
                pl = PyList_New(1);
                PyList_SET_ITEM(pl, 0, Py_BuildValue("(if)", r[20 + ri], 1.0f));
            }
            PyList_SET_ITEM(pols, ri, pl);
        }
        put("policies", pols);
        PyObject* fs = PyList_New(R);
        PyObject* ps = PyList_New(R);
        for (int ri = 0; ri < R; ri++) {
            PyList_SET_ITEM(fs, ri, PyFloat_FromDouble(f[12 + ri]));
            PyList_SET_ITEM(ps, ri, PyFloat_FromDouble(f[16 + ri]));
        }
        put("final_score", fs);
        put("depth", PyLong_FromLong(r[1]));
        put("game_length", PyLong_FromLong(r[2]));
        std::string id = *self->identifier + (many ? "_m" + std::to_string(man[i]) : "") + "_" + std::to_string(r[3]) + "_" + std::to_string(r[4]);
        put("match_identifier", PyUnicode_FromString(id.c_str()));
        put("has_resigned", PyBool_FromLong(r[5]));
        put("resign_false_positive", PyBool_FromLong(r[6]));
        put("starting_sample_depth", PyLong_FromLong(r[7]));
        put("resultant_puct_score", ps);
        put("resultant_puct_visits", PyLong_FromLong(r[8]));
        PyList_SET_ITEM(out, i, d);
    }
    return out;
}
static PyObject* supervisor_add_unique_state(PySupervisor*, PyObject* args) {
    Py_ssize_t p = 0;
    if (!PyArg_ParseTuple(args, "n", &p)) return nullptr;
    Py_RETURN_NONE;             // the reference's Python wrapper never calls it (cppinterface.py:218-219)
}
static PyObject* supervisor_clear_unique_states(PySupervisor* self, PyObject*) {
    if (gz_supervisor_clear_unique_states(self->s) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* supervisor_set_num_workers(PySupervisor*, PyObject*) { Py_RETURN_NONE; }
static PyObject* supervisor_totals(PySupervisor* self, PyObject*) {
    double t[26];
    if (gz_supervisor_totals(self->s, t) != GZ_OK) return fail();
    static const char* K[26] = {"evals", "moves", "games", "samples", "playouts", "errors", "halted", "first_error", "max_nodes",
        "max_children", "dupes", "no_samples", "fp_resign0", "fp_resign1", "early_run_to_ends", "resigns", "aborts",
        "waiting_eval", "ready", "select_nodes", "select_children", "new_children", "depth_sum", "dropped_samples", "error_restarts",
        "transpositions"};
    PyObject* d = PyDict_New();
    for (int i = 0; i < 26; i++) { PyObject* v = PyFloat_FromDouble(t[i]); PyDict_SetItemString(d, K[i], v); Py_DECREF(v); }
    return d;
}
static PyMethodDef supervisor_methods[] = {
    {"start_self_play", (PyCFunction)supervisor_start_self_play, METH_VARARGS, "start_self_play(num_workers, conf_dict)"},
    {"poll", (PyCFunction)supervisor_poll, METH_VARARGS, "poll(n_predictions, [policy_0.., value]) -> planes or None"},
    {"fetch_samples", (PyCFunction)supervisor_fetch_samples, METH_NOARGS, "finished samples as dicts, or None"},
    {"add_unique_state", (PyCFunction)supervisor_add_unique_state, METH_VARARGS, "add_unique_state(bs_ptr)"},
    {"clear_unique_states", (PyCFunction)supervisor_clear_unique_states, METH_NOARGS, "clear_unique_states()"},
    {"set_num_workers", (PyCFunction)supervisor_set_num_workers, METH_VARARGS, "accepted and ignored (cppinterface.py:197-198)"},
    {"totals", (PyCFunction)supervisor_totals, METH_NOARGS, "(additive) the managers' counters"},
    {nullptr, nullptr, 0, nullptr}};
static PyTypeObject PySupervisorType = {PyVarObject_HEAD_INIT(nullptr, 0)};

static PyObject* make_supervisor(PyObject*, PyObject* args) {
    Py_ssize_t sm_ptr = 0;
    PyObject* tr = nullptr;
    int batch = 0;
    const char* ident = nullptr;
    if (!PyArg_ParseTuple(args, "nO!is", &sm_ptr, &PyTransformerType, &tr, &batch, &ident)) return nullptr;
    gz_propnet* pn = (gz_propnet*)sm_ptr;
    gz_transformer* t = ((PyTransformer*)tr)->t;
    if (gz_transformer_finalize(t, (int)gz_propnet_info(pn, GZ_INFO_N_BASES), 0) != GZ_OK) return fail();
    gz_supervisor* s = nullptr;
    if (gz_supervisor_create(pn, t, batch, ident, 0, 0x5EED0000ull, 0, 0, &s) != GZ_OK) return fail();
    PySupervisor* o = PyObject_New(PySupervisor, &PySupervisorType);
    if (!o) { gz_supervisor_destroy(s); return nullptr; }
    o->s = s; o->pn = pn; o->transformer = tr; Py_INCREF(tr);
    o->identifier = new std::string(ident);
    return (PyObject*)o;
}

// ------------------------------------------------------------------------------------------------ player
struct PyPlayer {
    PyObject_HEAD
    gz_player* p;
    PyObject* transformer;
};
static void player_dealloc(PyPlayer* self) {
    if (self->p) gz_player_destroy(self->p);
    Py_XDECREF(self->transformer);
    Py_TYPE(self)->tp_free((PyObject*)self);
}
static PyObject* player_reset(PyPlayer* self, PyObject* args) {
    int depth = 0;
    if (!PyArg_ParseTuple(args, "i", &depth)) return nullptr;
    if (gz_player_reset(self->p, depth) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* player_apply_move(PyPlayer* self, PyObject* args) {
    Py_ssize_t jm = 0;
    if (!PyArg_ParseTuple(args, "n", &jm)) return nullptr;
    if (gz_player_apply_move(self->p, (const int*)jm, (int)gz_player_info(self->p, 1)) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* player_move(PyPlayer* self, PyObject* args) {
    Py_ssize_t bs = 0;
    int evals = 0;
    double end_time = 0;
    if (!PyArg_ParseTuple(args, "nid", &bs, &evals, &end_time)) return nullptr;
    if (gz_player_move(self->p, (const uint32_t*)bs, evals, end_time) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* player_get_move(PyPlayer* self, PyObject* args) {
    int role = 0, move = -1, nodes = -1;
    float score = -1;
    if (!PyArg_ParseTuple(args, "i", &role)) return nullptr;
    if (gz_player_get_move(self->p, role, &move, &score, &nodes) != GZ_OK) return fail();
    return Py_BuildValue("(ifi)", move, score, nodes);
}
static PyObject* player_update_config(PyPlayer* self, PyObject* args) {
    double think = 0;
    int conv = 0, verbose = 0;
    if (!PyArg_ParseTuple(args, "dii", &think, &conv, &verbose)) return nullptr;
    if (gz_player_update_config(self->p, think, conv, verbose) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* player_balance_moves(PyPlayer* self, PyObject* args) {
    int n = 0;
    if (!PyArg_ParseTuple(args, "i", &n)) return nullptr;
    if (gz_player_balance_moves(self->p, n) != GZ_OK) return fail();
    Py_RETURN_NONE;
}
static PyObject* player_tree_debug(PyPlayer* self, PyObject* args) {
    int n = 0;
    if (!PyArg_ParseTuple(args, "i", &n)) return nullptr;
    const int m = n > 0 ? n : 1;
    std::vector<float> sc(m);
    std::vector<int> ld(m), mv(m), vl(m), vp((size_t)m * 20);
    const int k = gz_player_tree_debug_lines(self->p, n > 0 ? n : 0, sc.data(), ld.data(), mv.data(), vl.data(), vp.data());
    if (k < 0) return fail();
    PyObject* out = PyTuple_New(k);
    for (int i = 0; i < k; i++) {        // (score, lead role, move, ((lead role, move), ..)) as player_impl.cpp:75-100
        PyObject* var = PyTuple_New(vl[i]);
        for (int j = 0; j < vl[i]; j++) PyTuple_SET_ITEM(var, j, Py_BuildValue("(ii)", vp[i * 20 + 2 * j], vp[i * 20 + 2 * j + 1]));
        PyTuple_SET_ITEM(out, i, Py_BuildValue("(fiiN)", sc[i], ld[i], mv[i], var));
    }
    return out;
}
static PyObject* player_poll(PyPlayer* self, PyObject* args) {
    int n = 0;
    PyObject* list = nullptr;
    if (!PyArg_ParseTuple(args, "iO!", &n, &PyList_Type, &list)) return nullptr;
    std::vector<const float*> ptrs;
    if (n > 0 && !prediction_pointers(list, ptrs)) return nullptr;
    const float* planes = nullptr;
    const int rows = gz_player_poll(self->p, n, ptrs.empty() ? nullptr : ptrs.data(), (int)ptrs.size(), &planes);
    if (rows < 0) return fail();
    if (rows == 0) Py_RETURN_NONE;
    return planes_array(planes, (Py_ssize_t)rows * gz_player_info(self->p, 0));     // rows leaves, flat (common.cpp:200-218)
}
static PyMethodDef player_methods[] = {
    {"player_reset", (PyCFunction)player_reset, METH_VARARGS, "player_reset(game_depth)"},
    {"player_apply_move", (PyCFunction)player_apply_move, METH_VARARGS, "player_apply_move(joint_move_ptr)"},
    {"player_move", (PyCFunction)player_move, METH_VARARGS, "player_move(basestate_ptr, evaluations, end_time)"},
    {"player_get_move", (PyCFunction)player_get_move, METH_VARARGS, "player_get_move(role) -> (legal index, score, nodes)"},
    {"player_update_config", (PyCFunction)player_update_config, METH_VARARGS, "player_update_config(think_time, converged_visits, verbose)"},
    {"player_balance_moves", (PyCFunction)player_balance_moves, METH_VARARGS, "player_balance_moves(max_count)"},
    {"player_tree_debug", (PyCFunction)player_tree_debug, METH_VARARGS, "player_tree_debug(max_count) -> ((score, lead, move, variation), ..)"},
    {"poll", (PyCFunction)player_poll, METH_VARARGS, "poll(n_predictions, [policy_0.., value]) -> planes or None"},
    {nullptr, nullptr, 0, nullptr}};
static PyTypeObject PyPlayerType = {PyVarObject_HEAD_INIT(nullptr, 0)};

static PyObject* make_player(PyObject*, PyObject* args) {
    Py_ssize_t sm_ptr = 0;
    PyObject *tr = nullptr, *conf = nullptr;
    if (!PyArg_ParseTuple(args, "nO!O!", &sm_ptr, &PyTransformerType, &tr, &PyDict_Type, &conf)) return nullptr;
    double pc[22];
    if (!puct_array(conf, pc)) return nullptr;
    gz_propnet* pn = (gz_propnet*)sm_ptr;
    gz_transformer* t = ((PyTransformer*)tr)->t;
    if (gz_transformer_finalize(t, (int)gz_propnet_info(pn, GZ_INFO_N_BASES), 0) != GZ_OK) return fail();
    gz_player* p = nullptr;
    if (gz_player_create(pn, t, pc, 0, 0x5EED0001ull, 0, 0, &p) != GZ_OK) return fail();
    PyPlayer* o = PyObject_New(PyPlayer, &PyPlayerType);
    if (!o) { gz_player_destroy(p); return nullptr; }
    o->p = p; o->transformer = tr; Py_INCREF(tr);
    return (PyObject*)o;
}

// ------------------------------------------------------------------------------------------------ module
// bytes -> tuple of bools, most significant bit first in every byte (ggpzero_interface.cpp:22-46)
static PyObject* buf_to_tuple_reverse_bytes(PyObject*, PyObject* args) {
    PyObject* o = nullptr;
    if (!PyArg_ParseTuple(args, "O", &o)) return nullptr;
    char* data = nullptr;
    Py_ssize_t n = 0;
    if (PyBytes_AsStringAndSize(o, &data, &n) < 0) return nullptr;
    PyObject* t = PyTuple_New(n * 8);
    for (Py_ssize_t i = 0; i < n; i++)
        for (int b = 0; b < 8; b++) PyTuple_SET_ITEM(t, i * 8 + b, PyBool_FromLong(((unsigned char)data[i] >> (7 - b)) & 1));
    return t;
}

static PyMethodDef module_methods[] = {
    {"GdlBasesTransformer", make_transformer, METH_VARARGS, "GdlBasesTransformer(channel_size, channels_per_state, num_control_channels, num_prev_states, num_rewards, [policy sizes])"},
    {"Supervisor", make_supervisor, METH_VARARGS, "Supervisor(sm_ptr, transformer, batch_size, identifier)"},
    {"Player", make_player, METH_VARARGS, "Player(sm_ptr, transformer, puct_conf_dict)"},
    {"buf_to_tuple_reverse_bytes", buf_to_tuple_reverse_bytes, METH_VARARGS, "bytes -> tuple of bools, MSB first"},
    {nullptr, nullptr, 0, nullptr}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "ggpzero_interface",
                                       "galvanise_zero's extension module on the B200-native engine (libgz_b200)", -1,
                                       module_methods, nullptr, nullptr, nullptr, nullptr};

static bool ready_type(PyTypeObject* t, const char* name, size_t size, destructor dealloc, PyMethodDef* methods) {
    t->tp_name = name;
    t->tp_basicsize = (Py_ssize_t)size;
    t->tp_dealloc = dealloc;
    t->tp_flags = Py_TPFLAGS_DEFAULT;
    t->tp_methods = methods;
    return PyType_Ready(t) == 0;
}

PyMODINIT_FUNC PyInit_ggpzero_interface(void) {
    import_array();
    if (!ready_type(&PyTransformerType, "ggpzero_interface.GdlBasesTransformerWrapper", sizeof(PyTransformer), (destructor)transformer_dealloc, transformer_methods) ||
        !ready_type(&PySupervisorType, "ggpzero_interface.SupervisorWrapper", sizeof(PySupervisor), (destructor)supervisor_dealloc, supervisor_methods) ||
        !ready_type(&PyPlayerType, "ggpzero_interface.PlayerWrapper", sizeof(PyPlayer), (destructor)player_dealloc, player_methods))
        return nullptr;
    PyObject* m = PyModule_Create(&moduledef);
    if (!m) return nullptr;
    GzError = PyErr_NewException("ggpzero_interface.GGPZeroInterfaceError", nullptr, nullptr);
    Py_INCREF(GzError);
    PyModule_AddObject(m, "GGPZeroInterfaceError", GzError);
    return m;
}
