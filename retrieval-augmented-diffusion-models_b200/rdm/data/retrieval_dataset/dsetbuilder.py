"""`rdm.data.retrieval_dataset.dsetbuilder.DatasetBuilder` -- the retrieval-database side of the reference
(`rdm/data/retrieval_dataset/dsetbuilder.py:49-655`) reduced to what sampling uses:

* `load_embeddings` (`:199-236`): `.npz` file or directory of parts with keys `embedding`, `img_id`, `patch_coords`;
* `train_searcher` (`:534-619`): the ScaNN build/load is replaced by uploading the rows to HBM once
  (`rdm_b200.knn.B200Searcher`: exact cosine top-k, per-row inverse norms computed on the device) -- no k-means,
  no quantiser training, nothing to serialise; row-sharded across ranks when torch.distributed is initialised and
  `shard=True`;
* `search_k_nearest` (`:478-518`) with the same result dictionary;
* `embed` (`:461-473`) through the retriever (CLIP).
* `build_data_pool` / `save_datapool` / `reset_data_pool` (`:238-262,317-437`): bulk CLIP-image embedding of caller-supplied patch batches
  and the chunked `.npz` pool writer (the image datasets themselves and visualisation stay out of scope).
"""
import datetime
import os
import time
from glob import glob

import numpy as np
import torch

from ldm.util import instantiate_from_config
from rdm_b200 import db_loader
from rdm_b200.knn import B200Searcher, ShardedSearcher, shard_range


class DatasetBuilder(object):
    def __init__(self, retriever_config, data=None, metric='dot_product', patch_size=128, n_patches=None, batch_size=10,
                 patch_sampling='random', k=10, img_size=None, num_workers=None, max_pool_size=None, visualize=False, save=True,
                 saved_embeddings=None, trainset_size_partitioning=None, chunk_size=None, gpu=True, load_patch_dataset=True,
                 patch_dset_kwargs=None, searcher_savepath=None, timestamp_searcher_savepath=False, savepath_postfix=None,
                 save_searcher=False, shard=False, direct_to_hbm=False):
        assert metric == 'dot_product', "the reference configs search by dot product on normalised vectors"
        self.retriever_config = retriever_config
        self.retriever_name = retriever_config["target"].split('.')[-1] if retriever_config else "none"
        self.visualize, self.distance_metric, self.k, self.chunk_size = visualize, metric, k, chunk_size
        self.timestamp = datetime.datetime.now().strftime("%Y-%m-%dT%H-%M-%S")
        self.load_patch_dataset = bool(load_patch_dataset)
        self.max_pool_size, self.patch_size, self.save_searcher = max_pool_size, patch_size, save_searcher
        self.dset = self.patch_dset = None
        if self.load_patch_dataset:
            print(f'WARNING: {self.__class__.__name__} (B200 build) does not load image patch datasets; neighbour images are unavailable')
            self.load_patch_dataset = False
        self.retriever_bs = batch_size
        self.gpu = gpu and torch.cuda.is_available()
        self._retriever = None                        # CLIP is instantiated lazily: sampling from DB rows never needs it
        self.data_pool = {'embedding': [], 'img_id': [], 'patch_coords': []}
        self.saved_embeddings = saved_embeddings
        self.shard = shard
        # direct_to_hbm (not a reference key): the embedding rows go from the .npz parts straight into the searcher's device buffer through
        # pinned staging chunks (rdm_b200.db_loader.load_rows_to_device) -- no host concatenation, no host copy of the rows at all;
        # data_pool['embedding'] is then a DeviceRows view (len / shape / fancy indexing by device gather)
        self.direct_to_hbm = bool(direct_to_hbm)
        self.load_stats = None
        self._row_base, self._n_total = None, None   # set when only this rank's rows of a sharded database were read
        self.searcher = None
        self.searcher_savedir = searcher_savepath
        self.savepath_postfix = savepath_postfix
        self.dset_name = data["target"].split('.')[-1] if data is not None and "target" in data else "dataset"
        if self.saved_embeddings:
            self.load_embeddings()

    @property
    def num_rows(self):
        """GLOBAL number of database rows: `len(data_pool['embedding'])` of the reference (ddpm.py:867), also when this rank holds only
        its row range of a sharded database (`get_qids` must draw pseudo-queries from the whole database on every rank)."""
        return int(self._n_total) if self._n_total is not None else len(self.data_pool['embedding'])

    # ---- retriever (CLIP) ------------------------------------------------------------------------------------
    @property
    def retriever(self):
        if self._retriever is None and self.retriever_config:
            self._retriever = self.load_retriever(gpu=self.gpu)
        return self._retriever

    def load_retriever(self, gpu=True, eval_mode=True):
        model = instantiate_from_config(self.retriever_config)
        if gpu and hasattr(model, "cuda"):
            model.cuda()
        if eval_mode and hasattr(model, "eval"):
            model.eval()
        return model

    @torch.no_grad()
    def embed(self, batch, is_caption=False):
        if not is_caption:
            if isinstance(batch, np.ndarray):
                batch = torch.from_numpy(batch)
            if batch.ndim == 5:
                batch = batch.reshape(-1, *batch.shape[2:])
            if batch.shape[-1] in (1, 3):
                batch = batch.permute(0, 3, 1, 2)
            batch = batch.contiguous().float()
            bs = batch.shape[0]
        else:
            bs = len(batch)
        return self.retriever(batch).reshape(bs, -1)

    # ---- database ------------------------------------------------------------------------------------------------
    def load_single_file(self, saved_embeddings):
        assert saved_embeddings.endswith('.npz'), 'saved embeddings not stored as a .npz file'
        compressed = np.load(saved_embeddings)
        self.data_pool = {key: compressed[key] for key in compressed.files}
        n = self.data_pool['embedding'].shape[0]
        if self.max_pool_size is None or n >= self.max_pool_size:
            self.max_pool_size = n
        print('Finished loading of patch embeddings.')

    def load_embeddings(self):
        if len(self.data_pool['embedding']) > 0:
            return
        print(f'Load saved patch embedding from "{self.saved_embeddings}"')
        if self.direct_to_hbm and (os.path.isfile(self.saved_embeddings) or os.path.isdir(self.saved_embeddings)):
            return self._load_direct()
        if self._dist_world() > 1 and (os.path.isfile(self.saved_embeddings) or os.path.isdir(self.saved_embeddings)):
            # row-sharded database: this rank reads only the parts that overlap the rows it will own; the small
            # id / coordinate arrays stay complete on every rank because search results index them globally (dsetbuilder.py:494-495)
            r, w = torch.distributed.get_rank(), torch.distributed.get_world_size()
            n_total = sum(db_loader.part_row_counts(db_loader.list_parts(self.saved_embeddings)))
            lo, hi = shard_range(n_total, r, w)
            self.data_pool = {'embedding': db_loader.load_rows(self.saved_embeddings, lo, hi, keys=('embedding',))['embedding']}
            meta = db_loader.load_rows(self.saved_embeddings, 0, n_total, keys=('img_id', 'patch_coords'))
            self.data_pool.update({k: meta[k] for k in ('img_id', 'patch_coords') if k in meta})
            self._row_base, self._n_total = lo, n_total
            if self.max_pool_size is None or n_total >= self.max_pool_size:
                self.max_pool_size = n_total
            print(f'Rank {r}/{w}: rows [{lo}, {hi}) of the {n_total}-row retrieval database.')
            return
        if os.path.isfile(self.saved_embeddings):
            self.load_single_file(self.saved_embeddings)
        elif os.path.isdir(self.saved_embeddings):
            files = sorted(glob(os.path.join(self.saved_embeddings, '*.npz')))
            if len(files) == 1:
                self.load_single_file(files[0])
            else:
                parts = [np.load(f) for f in files]
                keys = parts[0].files
                self.data_pool = {key: np.concatenate([p[key] for p in parts], axis=0) for key in keys}
        else:
            raise ValueError(f'Embeddings string "{self.saved_embeddings}" nor directory neither file --> check this.')
        print(f'Finished loading of retrieval database of length {self.data_pool["embedding"].shape[0]}.')

    def _load_direct(self):
        """f-3: this rank's rows (all rows without sharding) from disk to HBM in pinned-staged chunks; the searcher is ready afterwards."""
        n_total = sum(db_loader.part_row_counts(db_loader.list_parts(self.saved_embeddings)))
        w = self._dist_world()
        r = torch.distributed.get_rank() if w > 1 else 0
        lo, hi = shard_range(n_total, r, w)
        dev = torch.device("cuda", torch.cuda.current_device())
        rows, self.load_stats = db_loader.load_rows_to_device(self.saved_embeddings, lo, hi, dev)
        local = B200Searcher(rows, device=dev, idx_base=lo)
        self.searcher = ShardedSearcher(local) if w > 1 else local
        meta = db_loader.load_rows(self.saved_embeddings, 0, n_total, keys=('img_id', 'patch_coords'))
        self.data_pool = {'embedding': db_loader.DeviceRows(self.searcher, n_total, rows.shape[1], str(rows.dtype).split('.')[-1])}
        self.data_pool.update({k: meta[k] for k in ('img_id', 'patch_coords') if k in meta})
        self._row_base, self._n_total = (lo if w > 1 else None), n_total
        if self.max_pool_size is None or n_total >= self.max_pool_size:
            self.max_pool_size = n_total
        s = self.load_stats
        print(f'Rows [{lo}, {hi}) of {n_total} straight to {dev}: {s["bytes"] / 1e9:.2f} GB in {s["seconds"]:.2f} s ({s["gb_per_s"]:.2f} GB/s)')

    def _dist_world(self):
        """World size when the database is to be row-sharded (`shard=True` under an initialised torch.distributed), else 1."""
        d = torch.distributed
        return d.get_world_size() if (self.shard and d.is_available() and d.is_initialized()) else 1

    # ---- searcher ------------------------------------------------------------------------------------------------
    def train_searcher(self, k=None, metric=None, device=None, **ignored_scann_options):
        """Uploads the RAW rows to HBM (fp16 stays fp16) and computes the inverse norms there.  Replaces
        `scann.scann_ops_pybind.builder(emb / ||emb||, k, metric)...build()` (dsetbuilder.py:574-612)."""
        if self.searcher is not None:
            print('Using trained searcher')
            return
        emb = self.data_pool['embedding']
        assert len(emb) > 0, "no embeddings loaded"
        n, base = emb.shape[0], 0
        dist_on = self._dist_world() > 1
        if dist_on and self._row_base is not None:
            base = self._row_base                                        # load_embeddings already read just this rank's rows
        elif dist_on:
            r, w = torch.distributed.get_rank(), torch.distributed.get_world_size()
            base, end = shard_range(n, r, w)
            emb = emb[base:end]
        local = B200Searcher(emb, device=device, idx_base=base)
        self.searcher = ShardedSearcher(local) if dist_on else local
        print(f'Finish training searcher: {emb.shape[0]:,} rows resident on {local.device} (exact cosine top-k)')

    def search_k_nearest(self, queries, k=None, is_caption=False, visualize=None, query_embedded=False):
        assert self.searcher is not None, 'Cannot search with uninitialized searcher'
        k = self.k if k is None else k
        if not query_embedded:
            q_emb_ = self.embed(queries, is_caption=is_caption)
            q_emb_ = q_emb_.detach().cpu().numpy() if isinstance(q_emb_, torch.Tensor) else q_emb_
        else:
            q_emb_ = queries.detach().cpu().numpy() if isinstance(queries, torch.Tensor) else np.asarray(queries)
        query_embeddings = q_emb_ / np.linalg.norm(q_emb_, axis=1)[:, np.newaxis]
        start = time.time()
        nns, distances = self.searcher.search_batched(query_embeddings, final_num_neighbors=k)
        end = time.time()
        if self._row_base is not None:                                   # local rows only: gather across the shards on the device
            nn_emb = self.searcher.gather_device(torch.as_tensor(np.asarray(nns), dtype=torch.int64, device=self.searcher.local.device)).cpu().numpy()
        else:
            nn_emb = self.data_pool['embedding'][nns]
        out = {'embeddings': nn_emb, 'queries': queries, 'exec_time': end - start, 'nns': nns,
               'distances': distances, 'q_embeddings': q_emb_}
        for key_out, key in (('img_ids', 'img_id'), ('patch_coords', 'patch_coords')):
            if key in self.data_pool and len(self.data_pool[key]) > 0:
                out[key_out] = self.data_pool[key][nns]
        if visualize if visualize is not None else self.visualize:
            raise NotImplementedError("neighbour image patches need the patch dataset, which is outside the sampling hot path")
        return out

    def get_nn_patches(self, batched_nns):
        raise NotImplementedError("neighbour image patches need the patch dataset, which is outside the sampling hot path")

    # ---- database construction (reference dsetbuilder.py:238-262,317-437) ---------------------------------------
    def reset_data_pool(self):
        self.data_pool = {key: [] for key in self.data_pool}

    def save_datapool(self, postfix: str = None):
        """`np.savez_compressed` of the collected pool into `<base>/<dir_identifier>/<rows>x<dim>[-postfix].npz` -- the layout
        `load_embeddings` / `db_loader` read back.  `<base>` is the reference's hard-coded directory (dsetbuilder.py:248) unless
        `RDM_RETRIEVAL_DATASETS_DIR` (or the `pool_dir` attribute) names another one."""
        print('Save embeddings...')
        shape = list(self.data_pool['embedding'][0].shape)
        shape[0] *= len(self.data_pool['embedding'])                  # the reference's file name: rows of the first batch x number of batches
        identifier = 'x'.join(str(s) for s in shape)
        if postfix:
            print(f'Adding postfix "{postfix}" to identifier')
            identifier = identifier + '-' + postfix
        base = getattr(self, 'pool_dir', None) or os.environ.get('RDM_RETRIEVAL_DATASETS_DIR') or '/export/compvis-nfs/group/datasets/retrieval_datasets'
        img_dir = os.path.join(base, self.dir_identifier)
        os.makedirs(img_dir, exist_ok=True)
        self.saved_embeddings = img_dir
        saved_embeddings = f'{img_dir}/{identifier}.npz'
        self.data_pool = {key: np.concatenate(self.data_pool[key]) for key in self.data_pool}
        np.savez_compressed(saved_embeddings, **self.data_pool)
        return saved_embeddings

    @property
    def dir_identifier(self):
        ident = '-'.join([self.timestamp, getattr(self, 'dset_name', 'dataset'), self.retriever_name, str(self.patch_size)])
        return ident + (f"-{self.savepath_postfix}" if getattr(self, 'savepath_postfix', None) else '')

    def build_data_pool(self, loader=None, save=True, pool_dtype=np.float16):
        """Bulk embedding of image patches into a retrieval database (reference `build_data_pool`, dsetbuilder.py:317-437): every batch
        of `loader` -- dictionaries with `patch` ([b, h, w, 3] or [b, n, h, w, 3] in [-1, 1]), `img_id`, `patch_coords` and optionally
        `class_id`, i.e. what the reference's `PatcherDataset` + `custom_collate` deliver -- goes through `embed` (CLIP image tower on the
        device, one launch sequence per batch) and is appended to the pool; the pool is written as `part_<i>` files every `chunk_size`
        rows (or as one file at the end) and extraction stops at `max_pool_size` rows.  The image datasets themselves (ImageNet /
        OpenImages readers, patch sampling) stay with the caller: `loader` is any iterable, e.g. a `DataLoader` over the caller's patch
        dataset.  Rows are stored as float16 like the published databases (the reference's CLIP runs in half precision on the GPU).
        A pool that was loaded from `saved_embeddings` and is shorter than `max_pool_size` is continued: the batches whose rows are
        already present are skipped (the loader must replay the same order), new parts are numbered after the existing ones."""
        if loader is None:
            raise NotImplementedError("build_data_pool needs a loader of patch batches: the image patch datasets are outside this build (SURVEY.md section 2)")
        assert self.max_pool_size is not None, 'Max pool size still None --> check implementation'
        if self.chunk_size is not None:
            assert self.chunk_size % self.retriever_bs == 0, '"batch_size" has to evenly divide "chunk_size", if the latter is specified'
        n_examples = skip_rows = 0
        if self.saved_embeddings and len(self.data_pool['embedding']) > 0:
            current_len = self.data_pool['embedding'].shape[0]
            if current_len >= self.max_pool_size:
                print('embeddings are already saved, not recomputing....')
                return
            print(f'Restarting extraction as only {current_len} of overall {self.max_pool_size} examples are in data_pool.')
            n_examples = skip_rows = current_len
            self.data_pool = {key: [] for key in self.data_pool}
        self.data_pool = {key: ([] if not isinstance(v, list) else v) for key, v in self.data_pool.items()}
        part = int(n_examples / self.chunk_size) + 1 if self.chunk_size is not None else 1
        deltas, overall_start, seen, written = [], time.time(), 0, []

        def flush(postfix):
            if save and len(self.data_pool['embedding']) > 0:
                written.append(self.save_datapool(postfix=postfix))
                self.reset_data_pool()

        for batch in loader:
            if 'patch' not in batch:
                break
            patches = batch['patch']
            rows = int(np.prod(patches.shape[:-3]))
            if seen + rows <= skip_rows:                               # already in the saved pool
                seen += rows
                continue
            seen += rows
            start = time.time()
            embeddings = self.embed(patches)
            embeddings = (embeddings.detach().cpu().numpy() if isinstance(embeddings, torch.Tensor) else np.asarray(embeddings)).astype(pool_dtype, copy=False)
            deltas.append(time.time() - start)
            as_np = lambda v: v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            self.data_pool['patch_coords'].append(as_np(batch['patch_coords']))
            self.data_pool['img_id'].append(as_np(batch['img_id']))
            self.data_pool['embedding'].append(embeddings)
            if 'class_id' in batch:
                self.data_pool.setdefault('class_id', []).append(as_np(batch['class_id']))
            n_examples += embeddings.shape[0]
            if self.chunk_size is not None and n_examples / self.chunk_size >= part:
                flush(f'part_{part}')                                   # save in different chunks to avoid exceeding RAM
                part += 1
                if n_examples >= self.max_pool_size:
                    break
            elif self.chunk_size is None and n_examples >= self.max_pool_size:
                break
        if self.chunk_size is not None:
            flush(f'part_{part}')                                       # a last part smaller than chunk_size (nothing left after a complete chunk)
        elif save and len(self.data_pool['embedding']) > 0:             # only save a single file, when chunk size not defined
            written.append(self.save_datapool())
            self.reset_data_pool()
        overall, extract = time.time() - overall_start, float(np.sum(deltas))
        print(f'Finish extraction of {n_examples} feature embeddings')
        print('=' * 25, ' Time results ', '=' * 25)
        print(f'Extraction alone took {extract} secs; with loading {overall} secs; {extract / max(1, n_examples - skip_rows)} secs per sample')
        self.build_stats = {'rows': n_examples, 'new_rows': n_examples - skip_rows, 'embed_seconds': extract, 'seconds': overall, 'files': written}
        return written
