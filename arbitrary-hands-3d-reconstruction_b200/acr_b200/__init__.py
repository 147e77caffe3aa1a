"""acr_b200 -- B200-native runtime behind the drop-in ``acr`` / ``mano`` packages.

Layout of this directory (``arbitrary-hands-3d-reconstruction_b200/``):

* ``csrc/``      hand-written sm_100a CUDA kernels + the ``extern "C"`` boundary
                 (declared in ``/include/acr_b200.h``), built into ``lib/libacr_b200.so``.
* ``acr_b200/``  host runtime: ctypes binding, weight folding/packing, launch-plan
                 builder, synthetic assets, multi-GPU sharding.
* ``acr/``, ``mano/``  host-side mirror of the reference's Python call surface
                 (``acr.model.ACR``, ``acr.mano_wrapper.MANOWrapper``,
                 ``mano.manolayer.ManoLayer`` ...), so the directory can be put on
                 ``sys.path`` exactly like the reference's project root.

PyTorch is used for device memory, streams and ``torch.distributed`` only.
"""

__all__ = ["HOT_PATH_VERSION"]
HOT_PATH_VERSION = "r1"
