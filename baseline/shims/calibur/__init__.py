"""Shim: the three calibur functions rendering/camera.py:82-107 calls (standard GL perspective matrix)."""
import numpy


def fov_to_focal(fov, size):
    return size / (2.0 * numpy.tan(fov / 2.0))


def projection_gl_persp(width, height, cx, cy, fx, fy, near, far):
    return numpy.array([
        [2.0 * fx / width, 0.0, 1.0 - 2.0 * cx / width, 0.0],
        [0.0, 2.0 * fy / height, 2.0 * cy / height - 1.0, 0.0],
        [0.0, 0.0, (far + near) / (near - far), 2.0 * far * near / (near - far)],
        [0.0, 0.0, -1.0, 0.0]], dtype=numpy.float64)


def normalized(x):
    x = numpy.asarray(x, dtype=numpy.float64)
    return x / numpy.linalg.norm(x, axis=-1, keepdims=True)
