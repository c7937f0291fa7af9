"""Result store with the reference's layout (``HdfStore``, nestfit/main.py:233-377;
schema docs/store_spec.rst:58-150): ``<name>.store/`` holding ``table.hdf`` plus
one ``chunk<i>.hdf`` per worker, pixel groups ``/pix/<lon>/<lat>/<ncomp>`` with
the per-run attributes and datasets written by the dumper.

h5py / libhdf5 are optional.  When h5py imports, real HDF5 files with external
links are written exactly like the reference.  Otherwise the same tree is kept in
``MemGroup`` objects (a minimal h5py.Group look-alike) and persisted as
``.npz`` archives with the same file stems -- the API and schema are unchanged.
"""
import inspect
import json
import warnings
from pathlib import Path

import numpy as np

try:  # pragma: no cover - depends on the environment
    import h5py
    HAVE_H5PY = True
except Exception:  # h5py absent in this image
    h5py = None
    HAVE_H5PY = False


class _MemFile:
    def __init__(self, root):
        self.root = root

    def flush(self):
        pass


class MemGroup:
    """In-memory stand-in for ``h5py.Group``: ``attrs``, nested groups, datasets as
    numpy arrays, path addressing with '/'."""

    def __init__(self, name="/", parent=None):
        self.name = name
        self.attrs = {}
        self._items = {}
        self._parent = parent

    @property
    def file(self):
        return _MemFile(self._root())

    def _split(self, path):
        return [p for p in str(path).split("/") if p]

    def _root(self):
        g = self
        while g._parent is not None:
            g = g._parent
        return g

    def _resolve(self, path, create=False):
        """(parent group, leaf name) of `path`; absolute paths start at the root."""
        parts = self._split(path)
        g = self._root() if str(path).startswith("/") else self
        for part in parts[:-1]:
            nxt = g._items.get(part)
            if nxt is None:
                if not create:
                    raise KeyError(path)
                nxt = MemGroup(f"{g.name.rstrip('/')}/{part}", g)
                g._items[part] = nxt
            if not isinstance(nxt, MemGroup):
                raise KeyError(path)
            g = nxt
        return g, (parts[-1] if parts else None)

    def create_group(self, path):
        parent, leaf = self._resolve(path, create=True)
        if leaf in parent._items:
            raise ValueError(f"Unable to create group (name already exists): {path}")
        g = MemGroup(f"{parent.name.rstrip('/')}/{leaf}", parent)
        parent._items[leaf] = g
        return g

    def require_group(self, path):
        parent, leaf = self._resolve(path, create=True)
        if leaf is None:
            return parent
        g = parent._items.get(leaf)
        if g is None:
            g = MemGroup(f"{parent.name.rstrip('/')}/{leaf}", parent)
            parent._items[leaf] = g
        return g

    def create_dataset(self, name, data=None):
        parent, leaf = self._resolve(name, create=True)
        if leaf in parent._items:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        arr = np.array(data)
        parent._items[leaf] = arr
        return arr

    def __getitem__(self, path):
        parent, leaf = self._resolve(path)
        if leaf is None:
            return parent
        return parent._items[leaf]

    def __setitem__(self, path, value):
        parent, leaf = self._resolve(path, create=True)
        if isinstance(value, MemGroup):
            value._parent = parent
        parent._items[leaf] = value

    def __delitem__(self, path):
        parent, leaf = self._resolve(path)
        del parent._items[leaf]

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._items)

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    # ---- persistence --------------------------------------------------------
    def _flatten(self, prefix, out, attrs):
        if self.attrs:
            attrs[prefix or "/"] = {k: _jsonable(v) for k, v in self.attrs.items()}
        for k, v in self._items.items():
            p = f"{prefix}/{k}"
            if isinstance(v, MemGroup):
                v._flatten(p, out, attrs)
                if not v._items and not v.attrs:
                    attrs.setdefault(p, {})
            else:
                out[p] = v

    def save(self, path):
        out, attrs = {}, {}
        self._flatten("", out, attrs)
        out["__attrs__"] = np.array(json.dumps(attrs))
        with open(path, "wb") as f:
            np.savez_compressed(f, **out)

    @classmethod
    def load(cls, path):
        root = cls("/")
        with np.load(path, allow_pickle=False) as z:
            attrs = json.loads(str(z["__attrs__"]))
            for key in z.files:
                if key == "__attrs__":
                    continue
                root.create_dataset(key, data=z[key])
        for gpath, a in attrs.items():
            g = root.require_group(gpath)
            g.attrs.update({k: _unjson(v) for k, v in a.items()})
        return root


def _jsonable(v):
    if isinstance(v, np.ndarray):
        return {"__nd__": v.tolist()}
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    return v


def _unjson(v):
    if isinstance(v, dict) and "__nd__" in v:
        return np.array(v["__nd__"])
    return v


def check_ext(store_name, ext='hdf'):
    return store_name if store_name.endswith(f'.{ext}') else f'{store_name}.{ext}'


class HdfStore:
    """Store directory manager with the reference's names and methods
    (main.py:233-377).  `backend` is 'h5py' when available, else 'npz'."""
    linked_table = Path('table.hdf')
    chunk_prefix = 'chunk'
    dpath = '/products'

    def __init__(self, store_name, nchunks=1):
        self.store_name = str(store_name)
        self.store_dir = Path(check_ext(self.store_name, ext='store'))
        self.store_dir.mkdir(parents=True, exist_ok=True)
        self.backend = 'h5py' if HAVE_H5PY else 'npz'
        self._open = True
        if HAVE_H5PY:
            self.hdf = h5py.File(self.store_dir / self.linked_table, 'a')
        else:
            p = self._table_path
            self.hdf = MemGroup.load(p) if p.exists() else MemGroup("/")
        try:
            self.nchunks = int(self.hdf.attrs['nchunks'])
        except KeyError:
            self.hdf.attrs['nchunks'] = nchunks
            self.nchunks = nchunks
        try:
            from .models import MODELS
            self.model = MODELS[self.hdf.attrs['model_name']]
        except KeyError:
            self.model = None

    @property
    def _table_path(self):
        return self.store_dir / (str(self.linked_table) + ('' if HAVE_H5PY else '.npz'))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    @property
    def chunk_paths(self):
        suffix = '' if HAVE_H5PY else '.npz'
        return [self.store_dir / Path(f'{self.chunk_prefix}{i}.hdf{suffix}') for i in range(self.nchunks)]

    @property
    def is_open(self):
        return self._open

    def close(self):
        if not self._open:
            print('Store HDF already closed.')
            return
        if HAVE_H5PY:
            self.hdf.flush()
            self.hdf.close()
        else:
            self.hdf.save(self._table_path)
        self._open = False

    def open_chunk(self, i):
        """Group-like root of chunk `i` for a worker to write into."""
        path = self.chunk_paths[i]
        if HAVE_H5PY:
            return h5py.File(path, 'a')
        return MemGroup.load(path) if path.exists() else MemGroup("/")

    def close_chunk(self, i, root):
        if HAVE_H5PY:
            root.flush()
            root.close()
        else:
            root.save(self.chunk_paths[i])

    def iter_pix_groups(self):
        assert self.is_open
        for lon_pix in self.hdf['/pix']:
            for lat_pix in self.hdf[f'/pix/{lon_pix}']:
                group = self.hdf[f'/pix/{lon_pix}/{lat_pix}']
                if isinstance(group, np.ndarray):
                    continue
                yield group

    def find_first_valid_group(self):
        assert self.is_open
        for group in self.iter_pix_groups():
            if '1' in group:
                return group['1']
        raise ValueError('No valid pix groups found.')

    def link_files(self):
        """Make every chunk's /pix/<lon>/<lat> group reachable from the table
        (h5py: ExternalLinks like main.py:313-322; npz: mounted copies)."""
        assert self.is_open
        for chunk_path in self.chunk_paths:
            if not chunk_path.exists():
                continue
            if HAVE_H5PY:
                with h5py.File(chunk_path, 'r') as chunk_hdf:
                    if '/pix' not in chunk_hdf:
                        continue
                    for lon_pix in chunk_hdf['/pix']:
                        for lat_pix in chunk_hdf[f'/pix/{lon_pix}']:
                            name = f'/pix/{lon_pix}/{lat_pix}'
                            self.hdf[name] = h5py.ExternalLink(chunk_path.name, name)
                self.hdf.flush()
            else:
                chunk = MemGroup.load(chunk_path)
                if '/pix' not in chunk:
                    continue
                for lon_pix in chunk['/pix']:
                    for lat_pix in chunk[f'/pix/{lon_pix}']:
                        name = f'/pix/{lon_pix}/{lat_pix}'
                        self.hdf.require_group(f'/pix/{lon_pix}')
                        self.hdf[name] = chunk[name]

    def reset_pix_links(self):
        assert self.is_open
        if '/pix' in self.hdf:
            del self.hdf['/pix']

    def insert_header(self, stack):
        if self.is_open:
            sh_g = self.hdf.require_group('simple_header')
            for k, v in stack.simple_header.items():
                sh_g.attrs[k] = v
            fh_g = self.hdf.require_group('full_header')
            for k, v in stack.full_header.items():
                fh_g.attrs[k] = v
            self.hdf.attrs['naxis1'] = stack.shape[0]
            self.hdf.attrs['naxis2'] = stack.shape[1]
        else:
            warnings.warn('Could not insert header: the HDF5 file is closed.', category=RuntimeWarning)

    def read_header(self, full=True):
        """The stored cube header (`full`) or its two-dimensional map part (main.py:345-352): an
        ``astropy.io.fits.Header`` when astropy is importable, else a plain dict with the same cards."""
        assert self.is_open
        cards = dict(self.hdf['full_header' if full else 'simple_header'].attrs.items())
        try:
            from astropy.io import fits
        except ImportError:
            return cards
        header = fits.Header()                  # pragma: no cover - optional dependency
        for k, v in cards.items():              # pragma: no cover
            header[k] = v
        return header                           # pragma: no cover

    def create_dataset(self, dset_name, data, group='', clobber=True):
        assert len(dset_name) > 0
        g = self.hdf.require_group(group) if group else self.hdf
        if dset_name in g and clobber:
            warnings.warn(f'Deleting dataset "{group.rstrip("/")}/{dset_name}"', RuntimeWarning)
            del g[dset_name]
        return g.create_dataset(dset_name, data=data)

    def insert_fitter_pars(self, fitter):
        assert self.is_open
        self.hdf.attrs['lnZ_threshold'] = fitter.lnZ_thresh
        self.hdf.attrs['n_max_components'] = fitter.ncomp_max
        self.hdf.attrs['multinest_kwargs'] = str(fitter.mn_kwargs)

    def insert_model_metadata(self, runner_cls):
        module = inspect.getmodule(runner_cls)
        assert self.is_open
        self.hdf.attrs['n_params'] = module.N
        self.hdf.attrs['model_name'] = module.NAME
        self.hdf.attrs['par_names'] = module.PAR_NAMES
        self.hdf.attrs['par_names_short'] = module.PAR_NAMES_SHORT
        self.hdf.attrs['tex_labels'] = module.TEX_LABELS
        self.hdf.attrs['tex_labels_with_units'] = module.TEX_LABELS_WITH_UNITS
