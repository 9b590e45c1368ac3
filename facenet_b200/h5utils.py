"""``write_dict`` of the reference's ``facenet/h5utils.py:9-26`` -- the file format behind
``FaceToFaceValidation.write_h5file`` (statistics.py:330-331): every leaf of a (nested) dict becomes the dataset
``group/key/...`` of an HDF5 file, 1-D, resizable (``maxshape=(None,)``), gzip-compressed; writing the same name again
APPENDS to the dataset, so one file accumulates a value per validation.

``h5py`` is imported on use (the reference requires it; this image does not ship it).
"""
import numpy as np


def _leaves(dct, prefix=''):
    for key, item in dct.items():
        name = prefix + key
        if isinstance(item, dict):
            yield from _leaves(item, name + '/')
        else:
            yield name, np.atleast_1d(item)


def write_dict(file, dct, group=None):
    try:
        import h5py
    except ImportError as exc:
        raise ImportError('write_h5file / h5utils.write_dict need h5py') from exc
    prefix = group + '/' if group else ''
    with h5py.File(str(file), mode='a') as hf:
        for name, data in _leaves(dct, prefix):
            if name in hf:
                ds = hf[name]
                ds.resize(ds.shape[0] + data.shape[0], axis=0)
                ds[-data.shape[0]:] = data
            else:
                hf.create_dataset(name, data=data, maxshape=(None,), compression='gzip', dtype=data.dtype)
