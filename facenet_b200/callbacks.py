"""The caller side of the path: ``ValidateCallback.on_epoch_end`` (facenet/callbacks.py:21-28) and
``evaluate_embeddings`` (facenet/facenet.py:184-201) with the hand-off the GPU path wants.

The reference concatenates the model's outputs on the host (``np.concatenate`` forces a device -> host copy of every batch)
and ``FaceToFaceValidation`` would copy them back; here batches that already live on the GPU (torch tensors, TensorFlow
eager tensors through DLPack) are concatenated ON the device and handed to the validation in place.  Host batches take the
reference's route unchanged.  Framework-agnostic: ``ValidateCallback`` has the reference's constructor and ``on_epoch_end``;
subclass it together with ``tf.keras.callbacks.Callback`` when Keras drives the loop.
"""
import numpy as np

from facenet_b200 import statistics


def _to_device_tensor(batch):
    """torch view of a GPU batch (zero-copy), or None for host data"""
    if isinstance(batch, np.ndarray):
        return None
    if getattr(batch, 'is_cuda', False):
        return batch
    dev = getattr(batch, '__dlpack_device__', None)
    if dev is not None:
        try:
            if int(batch.__dlpack_device__()[0]) == 2:
                import torch
                return torch.from_dlpack(batch)          # TensorFlow eager tensor / CuPy array on the GPU
        except Exception:
            return None
    return None


def evaluate_embeddings(model, dset):
    """Evaluate embeddings for given data set (facenet/facenet.py:184-201): ``(embeddings [N, D], labels [N])``.  Embeddings stay
    on the GPU when the model produces them there; labels come back as a host array (the class bookkeeping is host work)."""
    embeddings_, labels_ = [], []
    on_gpu = None
    for images, labels in dset:
        emb = model(images)
        t = _to_device_tensor(emb)
        if on_gpu is None:
            on_gpu = t is not None
        embeddings_.append(t if on_gpu else np.asarray(emb))
        labels_.append(np.asarray(labels.cpu() if hasattr(labels, 'cpu') else labels))
    if not embeddings_:
        return np.zeros((0, 0), dtype=np.float32), np.zeros((0,), dtype=np.int64)
    if on_gpu:
        import torch
        return torch.cat([e.float() for e in embeddings_]).contiguous(), np.concatenate(labels_)
    return np.concatenate(embeddings_), np.concatenate(labels_)


class ValidateCallback:
    def __init__(self, model, dataset, every_n_epochs, max_nrof_epochs, config):
        self._model = model
        self.dataset = dataset
        self.config = config
        self.every_n_epochs = every_n_epochs
        self.max_nrof_epochs = max_nrof_epochs
        self.validation = None

    def on_epoch_end(self, epoch, logs=None):
        epoch1 = epoch + 1
        if epoch1 % self.every_n_epochs == 0 or epoch1 == self.max_nrof_epochs:
            embeddings, labels = evaluate_embeddings(self._model, self.dataset)
            self.validation = statistics.FaceToFaceValidation(embeddings, labels, self.config.validate)
