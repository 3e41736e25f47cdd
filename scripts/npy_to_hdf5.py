#!/usr/bin/env python
"""Pack the .npy model dump of cuNVSMTrainModel into the HDF5 file the reference writes.

The reference dumps `<output>_<epoch>.hdf5` with four float datasets named after `ModelBase::get_data()`
(cpp/model.cu:64-93, cpp/hdf5.cu:26-53: dims = {cols, rows}, i.e. row-major [objects, dim]) and its Python
tooling reads exactly those (py/nvsm/base.py:22-25,178-236). HDF5 is not part of this image, so the CLI writes
`<output>_<epoch>.<dataset>.npy` with the same names, shapes and dtype; this script — run where h5py exists —
produces the reference's file. `<output>_meta` (lse.Metadata) is already in the reference's format.

    python scripts/npy_to_hdf5.py <output> <epoch>      ->  <output>_<epoch>.hdf5
"""
import sys

import numpy as np

DATASETS = ("word_representations-representations", "entity_representations-representations",
            "word_entity_mapping-transform", "word_entity_mapping-bias")


def load_npy_dump(output, epoch):
    """The four tensors of a dump as {dataset name: float32 array [objects, dim]}."""
    return {name: np.load("%s_%s.%s.npy" % (output, epoch, name)) for name in DATASETS}


def main(argv):
    if len(argv) != 3:
        sys.stderr.write(__doc__)
        return 2
    try:
        import h5py
    except ImportError:
        sys.stderr.write("h5py is required to write HDF5 (it is not part of the build image)\n")
        return 1
    output, epoch = argv[1], argv[2]
    with h5py.File("%s_%s.hdf5" % (output, epoch), "x") as f:      # 'x': fail if the file exists (H5F_ACC_EXCL)
        for name, value in load_npy_dump(output, epoch).items():
            f.create_dataset(name, data=np.ascontiguousarray(value, dtype="<f4"))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
