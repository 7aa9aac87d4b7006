"""B200-native randomized-SVD engine behind the API of AMSC22-23/rSVD_Kamaneh_Raganato_Terrana.

The product is librsvdb.so (hand-written sm_100a CUDA behind the C ABI of include/rsvdb.h); this package holds its build
script, the ctypes binding and a Python mirror of the reference's operator interface.  Importing the package does not
load the library; the first use does, and fails loudly if it has not been built.
"""
from .capi import RsvdbError  # noqa: F401
from .engine import Engine, Image, PCA, POD, SVD, SVDMethod, rSVD, intermediate_step, generateOmega  # noqa: F401
