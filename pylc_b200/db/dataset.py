"""
Tile dataset -- host-side mirror of PyLC's db/ package (reference db/dataset.py:19-174,
db/database.py:20-260, db/buffer.py:17-103) collapsed into one module.

Same public surface: MLPDataset(db_path=None, input_data=None, partition=None, shuffle=False),
.loader(batch_size, n_workers, drop_last) -> (DataLoader, n_batches), .get_meta(), .get_data(key),
.save(), .size; iteration yields (img f32 [ch,T,T], mask i64 [T,T]) like the reference Buffer
(buffer.py:62-63).  Tiles may live on the host (u8 ndarray) or on the GPU (u8 CUDA tensors produced
by the extraction kernels); `device_batches()` walks GPU-resident tiles without a host round trip.

File format: the reference's HDF5 layout (datasets `img` [N,ch,T,T] u8 and `mask` [N,T,T] u8,
gzip, attribute `meta` = json of the Parameters; database.py:216-235) when h5py is importable;
otherwise -- h5py is not in this image -- a `.npz` with the same three keys.  Reading accepts both.
"""
import json
import math
import os

import numpy as np
import torch

from ..config import Parameters, defaults

try:  # optional: absent in this image
    import h5py
except Exception:  # pragma: no cover
    h5py = None


def _meta_to_json(meta):
    def conv(v):
        if isinstance(v, (np.ndarray, torch.Tensor)):
            return v.tolist()
        if isinstance(v, (np.integer,)):
            return int(v)
        if isinstance(v, (np.floating,)):
            return float(v)
        raise TypeError(type(v))
    return json.dumps(vars(meta), default=conv)


class DB(object):
    """Partitioned view of an (img, mask, meta) tile store (reference database.py:20-150)."""

    def __init__(self, path=None, data=None, partition=None, clip=None):
        assert (path is not None or data is not None) and not (path is not None and data is not None), \
            "Database requires either a path or input data to load."
        self.path = path
        self.data = data
        self.partition = partition if partition is not None else (0., 1.)
        self.clip = clip if clip is not None else defaults.clip
        if data is None:
            assert os.path.exists(path), "Database path {} does not exist.".format(path)
            self.data = self._read(path)
        self.img_shape = tuple(self.data['img'].shape)
        self.mask_shape = tuple(self.data['mask'].shape)
        self.size = int(self.clip * len(self.data['img']))
        self.start = int(math.ceil(self.partition[0] * self.size))
        self.end = int(math.ceil(self.partition[1] * self.size))
        self.partition_size = self.end - self.start
        self.buffer_size = min(defaults.buffer_size, self.partition_size)

    def __len__(self):
        return self.size

    @staticmethod
    def _read(path):
        if path.endswith('.npz'):
            with np.load(path, allow_pickle=False) as f:
                meta = Parameters().update(json.loads(str(f['meta'])))
                return {'img': f['img'], 'mask': f['mask'], 'meta': meta}
        if h5py is None:
            print('Error loading database:\n\t{}\nh5py is not installed; use the .npz container.'.format(path))
            exit(1)
        with h5py.File(path, mode='r', libver='latest', swmr=True) as f:
            meta = defaults.update(json.loads(f.attrs.get('meta')))   # database.py:160-164
            return {'img': f['img'][()], 'mask': f['mask'][()], 'meta': meta}

    def init_worker(self, worker_id, n_workers):
        """Split the partition across DataLoader workers (reference database.py:129-149)."""
        per_worker = int(math.ceil(self.partition_size / float(n_workers)))
        self.start += worker_id * per_worker
        self.end = min(self.start + per_worker, self.end)
        self.start = self.end if self.end < self.start else self.start
        self.partition_size = self.end - self.start
        self.buffer_size = min(defaults.buffer_size, self.partition_size)

    def get_meta(self):
        return self.data['meta']

    def get_data(self, key):
        return self.data[key]

    def save(self, file_path):
        assert file_path is not None, "File path must be specified to save data to database."
        img, mask = self.data['img'], self.data['mask']
        if len(img) == 0 or len(mask) == 0:
            print('\n --- Note: Image or mask data is empty.\n')
        img = img.cpu().numpy() if torch.is_tensor(img) else np.asarray(img)
        mask = mask.cpu().numpy() if torch.is_tensor(mask) else np.asarray(mask)
        meta_json = _meta_to_json(self.data['meta'])
        print('\nSaving buffer to database ... ')
        if h5py is not None and not file_path.endswith('.npz'):
            print('\nCopying {} samples to:\n\t{}  '.format(len(img), file_path))
            with h5py.File(file_path, 'w') as f:
                f.create_dataset("img", img.shape, compression='gzip', chunks=True, data=img)
                f.create_dataset("mask", mask.shape, compression='gzip', chunks=True, data=mask)
                f.attrs['meta'] = meta_json
        else:
            file_path = os.path.splitext(file_path)[0] + '.npz'
            print('\nCopying {} samples to:\n\t{}  '.format(len(img), file_path))
            np.savez_compressed(file_path, img=img, mask=mask, meta=np.array(meta_json))
        print('File saved.')
        return file_path


class MLPDataset(torch.utils.data.IterableDataset):
    def __init__(self, db_path=None, input_data=None, partition=None, shuffle=False):
        super().__init__()
        self.db = DB(path=db_path, data=input_data, partition=partition)
        self.shuffle = shuffle
        self.size = self.db.partition_size

    def _host_chunks(self):
        db = self.db
        for lo in range(db.start, db.end, max(db.buffer_size, 1)):
            hi = min(lo + db.buffer_size, db.end)
            imgs, masks = db.data['img'][lo:hi], db.data['mask'][lo:hi]
            if torch.is_tensor(imgs):
                imgs, masks = imgs.cpu().numpy(), masks.cpu().numpy()
            if self.shuffle:
                idx = np.random.permutation(hi - lo)
                imgs, masks = imgs[idx], masks[idx]
            yield imgs, masks

    def __iter__(self):
        """(img f32 [ch,T,T], mask i64 [T,T]) per tile, in buffer-sized chunks (buffer.py:47-65)."""
        info = torch.utils.data.get_worker_info()
        if info is not None and not getattr(self, "_worker_ready", False):
            self.db.init_worker(info.id, info.num_workers)
            self._worker_ready = True
        for imgs, masks in self._host_chunks():
            for i in range(len(imgs)):
                yield torch.tensor(imgs[i]).float(), torch.tensor(masks[i]).long()

    def loader(self, batch_size=1, n_workers=0, drop_last=False):
        return (torch.utils.data.DataLoader(self, batch_size=batch_size, num_workers=n_workers,
                                            worker_init_fn=self.init_worker,
                                            pin_memory=torch.cuda.is_available(), drop_last=drop_last),
                self.size // batch_size)

    def init_worker(self, worker_id):
        """DataLoader worker hook: each worker process iterates its own slice of the partition
        (reference dataset.py:109-115)."""
        info = torch.utils.data.get_worker_info()
        if info is not None and not getattr(self, "_worker_ready", False):
            self.db.init_worker(worker_id, info.num_workers)
            self._worker_ready = True

    def device_batches(self, batch_size, device=None, drop_last=False):
        """u8 (img [b,ch,T,T], mask [b,T,T]) batches as CUDA tensors.  GPU-resident tiles are sliced
        in place; host tiles are staged through pinned memory."""
        db = self.db
        device = device or torch.device("cuda", torch.cuda.current_device())
        for lo in range(db.start, db.end, batch_size):
            hi = min(lo + batch_size, db.end)
            if drop_last and hi - lo < batch_size:
                return
            imgs, masks = db.data['img'][lo:hi], db.data['mask'][lo:hi]
            if not torch.is_tensor(imgs):
                imgs = torch.from_numpy(np.ascontiguousarray(imgs)).pin_memory()
                masks = torch.from_numpy(np.ascontiguousarray(masks)).pin_memory()
            yield imgs.to(device, non_blocking=True), masks.to(device, non_blocking=True)

    def get_meta(self):
        return self.db.get_meta()

    def get_data(self, dset_key):
        return self.db.get_data(dset_key)

    def save(self):
        """Write `<id>.h5` (or `.npz`) into meta.output_dir or the default db dir (dataset.py:139-151)."""
        meta = self.get_meta()
        save_dir = meta.output_dir
        save_dir = save_dir if save_dir is not defaults.output_dir and os.path.isdir(save_dir) else defaults.db_dir
        os.makedirs(save_dir, exist_ok=True)
        return self.db.save(os.path.join(save_dir, meta.id + '.h5'))

    def print_meta(self, label=None):
        meta = self.get_meta()
        hline = '-' * 40
        print('\nDataset Configuration')
        print(hline)
        print('{:30s} {}'.format('Label', label if label is not None else '-'))
        print('{:30s} {}'.format('Database ID', meta.id))
        print('{:30s} {} ({})'.format('Channels', meta.ch, 'Grayscale' if meta.ch == 1 else 'Colour'))
        print('{:30s} {}px x {}px'.format('Tile size (WxH)', meta.tile_size, meta.tile_size))
        print('{:30s} {}'.format('Dataset Size', self.size))
        print('{:30s} {}'.format('Database Size', self.db.size))
        print('{:30s} {}'.format('Partition', self.db.partition))
        print('{:30s} {}'.format('Buffer Size', self.db.buffer_size))
        print(hline)
