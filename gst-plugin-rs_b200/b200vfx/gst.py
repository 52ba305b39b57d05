"""ctypes binding of the element layer (include/b200gst.h) -- lets the tests drive the elements the
way the reference's tests drive GStreamer: factory_make, set_property, start, push frames, read the bus."""
from __future__ import annotations

import ctypes as C

from . import FMT, lib as _vfx_lib, _ptr

FLOW_OK, FLOW_NOT_NEGOTIATED, FLOW_ERROR, FLOW_EOS = 0, -1, -2, -3
PAD_SINK, PAD_SRC = 0, 1
FMT_NAME = {v: k for k, v in FMT.items()}


class VideoFrameC(C.Structure):
    _fields_ = [("format", C.c_int), ("width", C.c_int), ("height", C.c_int), ("n_planes", C.c_int),
                ("data", C.c_void_p * 4), ("stride", C.c_int * 4)]


def frame(fmt: str, width: int, height: int, planes, strides) -> VideoFrameC:
    """planes: list of numpy arrays / tensors / addresses; strides: list of ints"""
    f = VideoFrameC()
    f.format, f.width, f.height, f.n_planes = FMT[fmt], width, height, len(planes)
    for i, (p, s) in enumerate(zip(planes, strides)):
        f.data[i] = _ptr(p)
        f.stride[i] = s
    return f


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _vfx_lib()
        vp, ci, cs = C.c_void_p, C.c_int, C.c_char_p
        fp = C.POINTER(VideoFrameC)
        ip = C.POINTER(C.c_int)
        sigs = {
            "b200gst_element_factory_make": ([cs], vp), "b200gst_element_unref": ([vp], None),
            "b200gst_element_factory_name": ([vp], cs), "b200gst_element_type_name": ([vp], cs),
            "b200gst_element_plugin_name": ([vp], cs), "b200gst_element_last_error": ([vp], cs),
            "b200gst_element_set_property": ([vp, cs, cs], ci), "b200gst_element_get_property": ([vp, cs, cs, C.c_size_t], ci),
            "b200gst_element_list_properties": ([vp, cs, C.c_size_t], ci),
            "b200gst_element_pad_template_formats": ([vp, ci, ip, ci], ci),
            "b200gst_element_transform_caps": ([vp, ci, ip, ci, ip, ci], ci),
            "b200gst_element_start": ([vp], ci), "b200gst_element_stop": ([vp], ci),
            "b200gst_element_set_caps": ([vp, ci, ci, ci, ci], ci), "b200gst_element_is_passthrough": ([vp], ci),
            "b200gst_element_transform_frame": ([vp, fp, fp], ci), "b200gst_element_transform_frame_ip": ([vp, fp], ci),
            "b200gst_roundedcorners_prepare_output": ([vp, fp, fp], ci),
            "b200gst_videocompare_request_pad": ([vp], ci), "b200gst_videocompare_release_pad": ([vp, ci], ci),
            "b200gst_videocompare_reference_pad": ([vp], ci),
            "b200gst_videocompare_aggregate_frames": ([vp, fp, ip, ci, C.c_int64, fp], ci),
            "b200gst_element_pop_message": ([vp, cs, C.c_size_t], ci),
        }
        for n, (a, r) in sigs.items():
            fn = getattr(L, n)
            fn.argtypes, fn.restype = a, r
        _lib = L
    return _lib


class Element:
    def __init__(self, factory: str):
        self._h = lib().b200gst_element_factory_make(factory.encode())
        if not self._h:
            raise ValueError("no such element factory: %s" % factory)

    def __del__(self):
        try:
            if self._h:
                lib().b200gst_element_unref(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def type_name(self):
        return lib().b200gst_element_type_name(self._h).decode()

    @property
    def plugin_name(self):
        return lib().b200gst_element_plugin_name(self._h).decode()

    @property
    def last_error(self):
        return lib().b200gst_element_last_error(self._h).decode()

    def set_property(self, name, value) -> int:
        v = None if value is None else str(value).encode()
        return lib().b200gst_element_set_property(self._h, name.encode(), v)

    def get_property(self, name):
        buf = C.create_string_buffer(1024)
        if lib().b200gst_element_get_property(self._h, name.encode(), buf, 1024) < 0:
            raise KeyError(name)
        return buf.value.decode()

    def list_properties(self):
        buf = C.create_string_buffer(4096)
        lib().b200gst_element_list_properties(self._h, buf, 4096)
        return [tuple(l.split("|")) for l in buf.value.decode().splitlines()]

    def pad_template_formats(self, direction):
        out = (C.c_int * 16)()
        n = lib().b200gst_element_pad_template_formats(self._h, direction, out, 16)
        return [FMT_NAME[out[i]] for i in range(n)]

    def transform_caps(self, direction, formats):
        arr = (C.c_int * len(formats))(*[FMT[f] for f in formats])
        out = (C.c_int * 16)()
        n = lib().b200gst_element_transform_caps(self._h, direction, arr, len(formats), out, 16)
        return [FMT_NAME[out[i]] for i in range(n)]

    def start(self):
        return lib().b200gst_element_start(self._h)

    def stop(self):
        return lib().b200gst_element_stop(self._h)

    def set_caps(self, in_fmt, out_fmt, w, h):
        return lib().b200gst_element_set_caps(self._h, FMT[in_fmt], FMT[out_fmt], w, h)

    @property
    def passthrough(self):
        return bool(lib().b200gst_element_is_passthrough(self._h))

    def transform_frame(self, fin, fout):
        return lib().b200gst_element_transform_frame(self._h, C.byref(fin), C.byref(fout))

    def transform_frame_ip(self, f):
        return lib().b200gst_element_transform_frame_ip(self._h, C.byref(f))

    def prepare_output(self, fin):
        out = VideoFrameC()
        rc = lib().b200gst_roundedcorners_prepare_output(self._h, C.byref(fin), C.byref(out))
        return rc, out

    def request_pad(self):
        return lib().b200gst_videocompare_request_pad(self._h)

    def release_pad(self, pad):
        return lib().b200gst_videocompare_release_pad(self._h, pad)

    @property
    def reference_pad(self):
        return lib().b200gst_videocompare_reference_pad(self._h)

    def aggregate_frames(self, frames, pad_ids, running_time_ns=-1, out=None):
        arr = (VideoFrameC * len(frames))(*frames)
        ids = (C.c_int * len(pad_ids))(*pad_ids)
        return lib().b200gst_videocompare_aggregate_frames(self._h, arr, ids, len(frames), running_time_ns,
                                                           C.byref(out) if out is not None else None)

    def pop_message(self):
        buf = C.create_string_buffer(4096)
        return buf.value.decode() if lib().b200gst_element_pop_message(self._h, buf, 4096) else None
