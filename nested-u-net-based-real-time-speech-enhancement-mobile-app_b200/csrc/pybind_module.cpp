// Thin pybind11 layer over the C ABI of include/nunet_b200.h (BASELINE.json north_star: "Python host calling hand-written
// sm_100a CUDA through a thin pybind11 C-ABI layer").  Nothing is computed here: every function forwards to the extern "C"
// entry point of the same name in libnunet_b200.so, with device / host addresses passed as integers (torch's data_ptr(),
// numpy's ctypes.data) exactly as a C caller would pass pointers.  Blocking calls release the GIL.
#include <pybind11/pybind11.h>

#include <cstdint>
#include <string>

#include "../../include/nunet_b200.h"

namespace py = pybind11;
using uptr = std::uintptr_t;

template <typename T>
static T* P(uptr a) {
    return reinterpret_cast<T*>(a);
}
static nunet_engine* H(uptr h) { return reinterpret_cast<nunet_engine*>(h); }

PYBIND11_MODULE(_nunet_pybind, m) {
    m.doc() = "pybind11 binding of the nunet_b200 C ABI (include/nunet_b200.h)";
    m.attr("ABI_VERSION") = NUNET_ABI_VERSION;
    m.def("last_error", [] { return std::string(nunet_last_error()); });
    m.def("abi_version", &nunet_abi_version);
    m.def("num_frames", &nunet_num_frames);
    m.def("blob_validate", [](py::bytes blob, int variant) {
        const std::string b = blob;
        return nunet_blob_validate(b.data(), b.size(), variant);
    });
    // returns (rc, handle)
    m.def("create", [](int variant, int device, int max_frames, int max_streams, int ctfa_mode, int dc_mode, int stream_ctfa_history,
                       int chunk_frames, py::bytes blob) {
        nunet_config cfg{variant, device, max_frames, max_streams, ctfa_mode, dc_mode, stream_ctfa_history, chunk_frames};
        const std::string b = blob;
        nunet_engine* h = nullptr;
        int rc;
        {
            py::gil_scoped_release nogil;
            rc = nunet_create(&cfg, b.data(), b.size(), &h);
        }
        return py::make_tuple(rc, reinterpret_cast<uptr>(h));
    });
    m.def("destroy", [](uptr h) { nunet_destroy(H(h)); }, py::call_guard<py::gil_scoped_release>());
    m.def("forward_wav_dev", [](uptr h, uptr wav, int B, int n, uptr out_wav, uptr out_mag, uptr stream) {
        return nunet_forward_wav_dev(H(h), P<const float>(wav), B, n, P<float>(out_wav), P<float>(out_mag), P<void>(stream));
    });
    m.def("forward_wav_host", [](uptr h, uptr wav, int B, int n, uptr out_wav, uptr out_mag) {
        return nunet_forward_wav_host(H(h), P<const float>(wav), B, n, P<float>(out_wav), P<float>(out_mag));
    }, py::call_guard<py::gil_scoped_release>());
    m.def("forward_mag_dev", [](uptr h, uptr mag, int B, int T, uptr out_mag, uptr stream) {
        return nunet_forward_mag_dev(H(h), P<const float>(mag), B, T, P<float>(out_mag), P<void>(stream));
    });
    m.def("stream_reset", [](uptr h, int first, int count, uptr stream) { return nunet_stream_reset(H(h), first, count, P<void>(stream)); });
    m.def("stream_step_mag_dev", [](uptr h, uptr mag, int S, uptr out_mag, uptr stream) {
        return nunet_stream_step_mag_dev(H(h), P<const float>(mag), S, P<float>(out_mag), P<void>(stream));
    });
    m.def("stream_step_wav_dev", [](uptr h, uptr hop, int S, uptr out_hop, uptr out_mag, uptr stream) {
        return nunet_stream_step_wav_dev(H(h), P<const float>(hop), S, P<float>(out_hop), P<float>(out_mag), P<void>(stream));
    });
    m.def("stream_step_wav_host", [](uptr h, uptr hop, int S, uptr out_hop) {
        return nunet_stream_step_wav_host(H(h), P<const float>(hop), S, P<float>(out_hop));
    }, py::call_guard<py::gil_scoped_release>());
    m.def("state_count", [](uptr h) { return nunet_state_count(H(h)); });
    // returns (rc, name)
    m.def("state_name", [](uptr h, int index) {
        char buf[128] = {0};
        const int rc = nunet_state_name(H(h), index, buf, sizeof buf);
        return py::make_tuple(rc, std::string(buf));
    });
    m.def("state_numel", [](uptr h, const std::string& name) { return nunet_state_numel(H(h), name.c_str()); });
    m.def("state_export", [](uptr h, int sid, const std::string& name, uptr buf) { return nunet_state_export(H(h), sid, name.c_str(), P<float>(buf)); });
    m.def("state_import", [](uptr h, int sid, const std::string& name, uptr buf) {
        return nunet_state_import(H(h), sid, name.c_str(), P<const float>(buf));
    });
    m.def("state_generation", [](uptr h) { return nunet_state_generation(H(h)); });
    m.def("last_launch_count", [](uptr h) { return nunet_last_launch_count(H(h)); });
    m.def("profile_enable", [](uptr h, int on) { return nunet_profile_enable(H(h), on); });
    m.def("profile_count", [](uptr h) { return nunet_profile_count(H(h)); });
    // returns (rc, name, ms, algorithmic bytes)
    m.def("profile_entry", [](uptr h, int index) {
        char buf[160] = {0};
        float ms = 0.f;
        double nb = 0.0;
        const int rc = nunet_profile_entry(H(h), index, buf, sizeof buf, &ms, &nb);
        return py::make_tuple(rc, std::string(buf), ms, nb);
    });
    m.def("debug_read", [](uptr h, const std::string& name, uptr buf, long long cap) { return nunet_debug_read(H(h), name.c_str(), P<float>(buf), cap); });
}
