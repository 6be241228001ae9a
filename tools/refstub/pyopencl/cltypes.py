"""numpy dtypes matching pyopencl.cltypes (only the ones the reference touches)."""
import numpy as _np

float = _np.float32
uint = _np.uint32
int = _np.int32
uchar = _np.uint8


def _vec(scalar, names):
    return _np.dtype([(n, scalar) for n in names])


float2 = _vec(_np.float32, "xy")
float3 = _vec(_np.float32, "xyzw")  # OpenCL float3 occupies 16 bytes
float4 = _vec(_np.float32, "xyzw")
uchar4 = _vec(_np.uint8, "xyzw")
uint2 = _vec(_np.uint32, "xy")
uint4 = _vec(_np.uint32, "xyzw")
