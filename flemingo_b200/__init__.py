"""flemingo_b200 -- B200-native (sm_100a) placement / fitness hot path of ErillLab/FLEMINGO.

Public surface (host side mirrors the reference's Python objects for this path only):

  objects.OrganismObject / PssmObject / ShapeObject / ConnectorObject / PlacementObject
      same constructor signatures, `flatten`, `get_placement`, `get_binding_energies`,
      `get_M_and_SE_on_DNA_set`, `set_fitness` as /root/reference/src/objects/*.py
  multiplacement.calculate(...)           drop-in for `_multiplacement.calculate` (src/_interface.c:45)
  engine.Engine                           batched population x sequence-set placement and fitness
  dropin.install()                        routes the reference's own OrganismObject through the GPU path

There is no CPU fallback: everything that computes a placement goes through libflemingo_b200.so
(include/flemingo_b200.h) and raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
