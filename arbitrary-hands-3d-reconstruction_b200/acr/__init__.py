"""Drop-in mirror of the reference's ``acr`` package for the inference hot path
(/root/reference/acr): same module names, classes, call signatures and output dict schema,
backed by the sm_100a kernels of libacr_b200.so.  Rendering, visualisation and the CLI loops
of the reference are out of scope."""
