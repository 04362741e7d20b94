"""Offline neighbour precompute: the `search_nns` / `save_pkl` part of the reference's
`scripts/search_neighbors.py:355-450` over the device-resident searcher.

Per query batch: CLIP-embed the query patches (or captions) on the device -> q / ||q|| -> exact kNN -> per example ONE pickle
`embeddings/{k}_nns-img{id:09d}.p` = `{n_patches_per_side: {'embeddings', 'img_ids', 'patch_coords', 'nn_ids'}}` -- the layout
`QueryDataset.load_nns` reads back (`rdm/data/base.py:925-939`).  Without `save`, the neighbour-id histogram the reference uses to build
`nn_memory` (`{'nn_memory', 'id_count'}`, `ddpm.py:168-176`) is returned.  Dataset construction / image I/O stay out of scope: the caller
supplies any iterable of `{'patches': [b, n, h, w, c] in [-1, 1]}` or `{'caption': [str, ...]}` batches with a `batch_size` attribute."""
import os
import pickle

import numpy as np
import torch


def save_pkl(filepath, save_it, npatches_perside, corrupts, i, j, start_id, dset_batch_size):
    """`scripts/search_neighbors.py:355-379`: merge the entry of this patch grid into an existing pickle, or create the file."""
    if os.path.isfile(filepath):
        try:
            with open(filepath, 'rb') as f:
                old_one = pickle.load(f)
            old_one.update({npatches_perside: save_it[npatches_perside]})
            with open(filepath, 'wb') as f:
                pickle.dump(old_one, f, protocol=pickle.HIGHEST_PROTOCOL)
        except Exception as e:
            print(f'ERROR: {e.__class__.__name__} : ', e)
            if npatches_perside == 1:
                print(f'Overwriting id {start_id + i * dset_batch_size + j} as it is corrupt.')
                with open(filepath, 'wb') as f:
                    pickle.dump(save_it, f, protocol=pickle.HIGHEST_PROTOCOL)
            else:
                corrupts.add(start_id + i * dset_batch_size + j)
                print(f'Adding id {start_id + i * dset_batch_size + j} to corrupts.')
    else:
        with open(filepath, 'wb') as f:
            pickle.dump(save_it, f, protocol=pickle.HIGHEST_PROTOCOL)
    return corrupts


def search_nns(dataset_builder, qloader, device='cuda', mode='img', save=False, npatches_perside=None, base_savedir=None, nn_paths=None,
               corrupts=None, start_id=0, max_its=None):
    """`scripts/search_neighbors.py:381-450`, same arguments and return values."""
    assert dataset_builder.searcher is not None
    dset_batch_size = qloader.batch_size
    if save:
        assert base_savedir is not None and npatches_perside is not None
        assert os.path.isdir(os.path.join(base_savedir, 'embeddings'))
        nn_paths = {} if nn_paths is None else nn_paths
        corrupts = set() if corrupts is None else corrupts
    return_ids = {}
    for i, batch in enumerate(qloader):
        if max_its is not None and i >= max_its:
            break
        query = batch['patches'].to(device) if mode == 'img' else batch['caption']
        if isinstance(query, torch.Tensor):
            b, n = query.shape[:2]
            query = query.reshape(b * n, *query.shape[2:])                       # 'b n h w c -> (b n) h w c'
        else:
            b, n = len(query), 1
        results = dataset_builder.search_k_nearest(query, visualize=False, is_caption=mode == 'text')
        if save:
            results = {key: results[key].reshape(b, n, *results[key].shape[1:]) if isinstance(results[key], np.ndarray) else results[key] for key in results}
            for j in range(len(results['embeddings'])):
                filename = f'embeddings/{dataset_builder.k}_nns-img{start_id + i * dset_batch_size + j:09d}.p'
                save_it = {npatches_perside: {'embeddings': results['embeddings'][j], 'img_ids': results['img_ids'][j],
                                              'patch_coords': results['patch_coords'][j], 'nn_ids': results['nns'][j]}}
                corrupts = save_pkl(os.path.join(base_savedir, filename), save_it, npatches_perside, corrupts, i, j, start_id, dset_batch_size)
                nn_paths.update({start_id + i * dset_batch_size + j: filename})
        else:
            ids, counts = np.unique(results['nns'], return_counts=True)
            for id_, c in zip(ids, counts):
                return_ids[int(id_)] = return_ids.get(int(id_), 0) + int(c)
    return nn_paths if save else return_ids


def build_nn_memory(return_ids):
    """`{'nn_memory': ids sorted by how often they were retrieved, 'id_count': {id: count}}` -- what `ddpm.py:168-176` loads."""
    order = sorted(return_ids.items(), key=lambda kv: (-kv[1], kv[0]))
    return {'nn_memory': np.asarray([k for k, _ in order], dtype=np.int64), 'id_count': dict(return_ids)}
