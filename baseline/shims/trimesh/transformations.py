import numpy


def translation_matrix(d):
    m = numpy.identity(4)
    m[:3, 3] = numpy.asarray(d, dtype=numpy.float64)[:3]
    return m


def inverse_matrix(m):
    return numpy.linalg.inv(m)


def translation_from_matrix(m):
    return numpy.asarray(m)[:3, 3].copy()
