"""Drop-in replacement of the reference's public functions (DINCAE/__init__.py:47-51,
:53-117, :120-230, :234-294, :302-329, :704-811): same names, same positional and keyword
arguments, same netCDF output layout -- executed by the sm_100a kernels of this package.
"""
import os
import random
from datetime import datetime, timedelta
from math import ceil

import numpy as np

from . import ncio

__all__ = ["reconstruct", "load_gridded_nc", "data_generator", "reconstruct_gridded_nc"]

NEAREST_NEIGHBOR = 1        # tf.image.ResizeMethod.NEAREST_NEIGHBOR in TF 1.15

# Arithmetic of the convolutions (new capability; the reference is fp32 throughout):
#   "tf32" (default) tcgen05 kind::tf32 on planar-4 fp32 activations,
#   "bf16"           tcgen05 kind::f16 on planar-8 bf16 activations (BASELINE configs[3]),
#   "fp32"           FFMA kernels, bit-for-bit fp32.
# Set with set_precision() or the environment variable DINCAE_PRECISION; the reference's
# keyword list of `reconstruct` stays unchanged.
_precision = [None]


def set_precision(name):
    """Select the convolution arithmetic of subsequent `reconstruct*` calls (see above)."""
    if name not in ("tf32", "bf16", "fp32", None):
        raise ValueError("precision must be 'tf32', 'bf16' or 'fp32'")
    _precision[0] = name


_dropout_active = [False]


def set_dropout_active(active):
    """``True``: ``dropout_rate_train`` of `reconstruct` drops units after every dense layer while
    training (what the reference's tf.layers.dropout call was meant to do).  Default ``False`` =
    the reference's actual behaviour: the layer is inert because it is called without
    ``training=True`` (DINCAE/__init__.py:483)."""
    _dropout_active[0] = bool(active)


def get_precision():
    name = _precision[0] or os.environ.get("DINCAE_PRECISION", "tf32").lower()
    if name not in ("tf32", "bf16", "fp32"):
        raise ValueError("DINCAE_PRECISION must be tf32, bf16 or fp32 (got %r)" % name)
    return name


def identity(x):
    return x


def sinv(x, minx=1e-3):
    """"Safe inversion" of DINCAE/__init__.py:298-299 for numpy / torch inputs."""
    try:
        import torch
        if isinstance(x, torch.Tensor):
            return 1 / torch.clamp(x, min=minx)
    except ImportError:                 # pragma: no cover
        pass
    return 1 / np.maximum(x, minx)


def load_gridded_nc(fname, varname, minfrac=0.05):
    """Same contract as the reference (:53-117): returns
    ``lon, lat, time, data (masked), missing (bool), mask (bool)``."""
    lon, lat, time, data, mask = ncio.read_gridded(fname, varname)
    data = np.ma.asarray(data)
    if mask is not None:
        mask = np.asarray(mask) == 1
    else:
        print("compute mask for ", varname, ": sea point should have at least ",
              minfrac, " for valid data tought time")
        if data.mask is np.ma.nomask or np.isscalar(data.mask):
            mask = np.ones((data.shape[1], data.shape[2]), dtype=bool)
        else:
            mask = np.mean(~data.mask, axis=0) > minfrac
        print("mask: sea points ", np.sum(mask))
        print("mask: land points ", np.sum(~mask))
    print("varname ", varname, mask.shape)
    if data.mask is np.ma.nomask or np.isscalar(data.mask):
        missing = np.zeros(data.shape, dtype=bool)
    else:
        missing = data.mask
    print("data shape: ", data.shape)
    # same print as the reference (:113); np.ma's min / max copy the cube twice (0.33 s for the
    # 5266-frame cube), the where= reductions do not
    raw, valid = np.ma.getdata(data), ~missing
    if valid.any():
        print("data range: ", raw.min(where=valid, initial=np.inf), raw.max(where=valid, initial=-np.inf))
    else:
        print("data range: ", np.ma.masked, np.ma.masked)
    return lon, lat, time, data, missing, mask


_recent_sources = []


class FrameSource:
    """Callable with the reference's generator protocol (``datagen()`` yields
    ``(xin[H,W,nvar] f32, xtrue[H,W,2] f32)`` frame by frame, :196-227) that also carries
    its source arrays, so ``reconstruct`` can keep the cube in HBM and build batches on the
    GPU instead of pulling frames through Python."""

    def __init__(self, lon, lat, time, data_full, missing, train, ntime_win, obs_err_std,
                 jitter_std):
        self.lon, self.lat, self.time = lon, lat, time
        self.data_full, self.missing = list(data_full), list(missing)
        self.train = bool(train)
        self.ntime_win = int(ntime_win)
        self.obs_err_std = list(obs_err_std)
        self.jitter_std = list(jitter_std)
        self._cube = None

    def same_data(self, other):
        return (isinstance(other, FrameSource) and len(other.data_full) == len(self.data_full)
                and all(a is b for a, b in zip(self.data_full, other.data_full))
                and all(a is b for a, b in zip(self.missing, other.missing))
                and self.ntime_win == other.ntime_win and self.obs_err_std == other.obs_err_std)

    def cube(self):
        """Device cube of this source; sources built from the same arrays share one."""
        from .data import DeviceCube
        if self._cube is None:
            for other in _recent_sources:
                if other._cube is not None and self.same_data(other):
                    self._cube = other._cube
                    break
            else:
                self._cube = DeviceCube(self.lon, self.lat, self.time, self.data_full, self.missing,
                                        self.obs_err_std, self.ntime_win)
            _recent_sources.append(self)
            del _recent_sources[:-4]
        return self._cube

    @property
    def ntime(self):
        return self.data_full[0].shape[0]

    def __call__(self, chunk=16):
        """Frames in time order, produced by the GPU batch builder and copied back."""
        import torch
        from . import layout
        cube = self.cube()
        T, H, W = cube.T, cube.H, cube.W
        xin = torch.zeros(chunk, cube.nvar_planes, H, W, 4, device=cube.dev)
        xtrue = torch.zeros(chunk, H, W, 2, device=cube.dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=cube.dev)
        seed = random.getrandbits(63) if self.train else 0
        for start in range(0, T, chunk):
            n = min(chunk, T - start)
            fi = torch.arange(start, start + n, dtype=torch.int32, device=cube.dev)
            ci = None
            if self.train:
                ci = torch.tensor([random.randrange(0, T) for _ in range(n)], dtype=torch.int32,
                                  device=cube.dev)
            cube.build_batch(fi, ci, n, xin, xtrue, cnt, jitter_std=self.jitter_std, seed=seed)
            a = layout.p4_to_nhwc(xin[:n].cpu().numpy(), cube.nvar)
            t = xtrue[:n].cpu().numpy()
            for k in range(n):
                yield a[k], t[k]


def data_generator_list(lon, lat, time, data_full, missing, train=True, ntime_win=3,
                        obs_err_std=[1.], jitter_std=[0.05]):
    """(:132-230) -> ``datagen, nvar, ntime, meandata``; ``meandata`` of the first variable."""
    sz = data_full[0].shape
    print("sz ", sz)
    src = FrameSource(lon, lat, time, data_full, missing, train, ntime_win, obs_err_std, jitter_std)
    cube = src.cube()
    return src, cube.nvar, cube.T, cube.meandata[0]


def data_generator(lon, lat, time, data_full, missing, train=True, ntime_win=3, obs_err_std=1.,
                   jitter_std=0.05):
    """(:120-130).  Like the reference, the ``train`` argument is NOT forwarded (the list
    version is always called with train=True; quirk Q1 of SURVEY.md)."""
    return data_generator_list(lon, lat, time, [data_full], [missing], train=True,
                               ntime_win=ntime_win, obs_err_std=[obs_err_std],
                               jitter_std=[jitter_std])


def savesample(fname, m_rec, σ2_rec, meandata, lon, lat, e, ii, offset,
               transfun=(identity, identity)):
    """(:234-294) write one batch of the reconstruction; creates the file at ``ii == 0``."""
    # (masked entries of meandata are masked again by the writer: plain arithmetic here)
    md = np.ma.asarray(meandata)
    recdata = transfun[1](m_rec + (np.ma.getdata(md) if md.mask is np.ma.nomask else md.filled(0)))
    if transfun[1] not in (np.exp, identity):
        print("warning: sigma_rec is not transformed")
    sigma_rec = np.sqrt(σ2_rec)
    w = _open_writers.get(fname)
    if ii == 0 or w is None:
        if w is not None:
            w.close()
        w = ncio.ReconWriter(fname, lon, lat, meandata, append=(ii != 0 and os.path.isfile(fname)))
        _open_writers[fname] = w
    w.write(offset, np.asarray(recdata), np.asarray(sigma_rec))
    # the reference opens and closes the file per call (:279-294): the file must be complete after
    # every call here as well (record count in the header), whoever closes the writer later
    w.sync()


_open_writers = {}


def _close_all_writers():
    for name in list(_open_writers):
        _close_writer(name)


import atexit  # noqa: E402

atexit.register(_close_all_writers)


def _close_writer(fname):
    w = _open_writers.pop(fname, None)
    if w is not None:
        w.close()


_last_stamp = [None]


def _output_name(outdir):
    """``data-%Y-%m-%dT%H%M%S.nc`` (:674-675); bumped by whole seconds when two saves fall
    into the same second (a B200 epoch is shorter than the 1 s resolution of the name)."""
    now = datetime.now().replace(microsecond=0)
    if _last_stamp[0] is not None and now <= _last_stamp[0]:
        now = _last_stamp[0] + timedelta(seconds=1)
    _last_stamp[0] = now
    return os.path.join(outdir, "data-{}.nc".format(now.strftime("%Y-%m-%dT%H%M%S")))


def _dp_context():
    """(torch.distributed, rank, world) when the caller runs one process per GPU inside an
    initialised process group of more than one rank (e.g. ``torchrun run_DINCAE.py`` after
    ``init_process_group("nccl")`` and ``torch.cuda.set_device``), else (None, 0, 1)."""
    try:
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
            return tdist, tdist.get_rank(), tdist.get_world_size()
    except Exception:            # noqa: BLE001
        pass
    return None, 0, 1


def reconstruct(lon, lat, mask, meandata,
                train_datagen, train_len,
                test_datagen, test_len,
                outdir,
                resize_method=NEAREST_NEIGHBOR,
                epochs=1000,
                batch_size=30,
                save_each=10,
                save_model_each=500,
                skipconnections=[1, 2, 3, 4],
                dropout_rate_train=0.3,
                tensorboard=False,
                truth_uncertain=False,
                shuffle_buffer_size=3 * 15,
                nvar=10,
                enc_nfilter_internal=[16, 24, 36, 54],
                frac_dense_layer=[0.2],
                clip_grad=5.0,
                regularization_L2_beta=0,
                transfun=(identity, identity),
                savesample=savesample,
                learning_rate=1e-3,
                learning_rate_decay_epoch=np.inf,
                iseed=None,
                nprefetch=0,
                loss=[],
                nepoch_keep_missing=0,
                ):
    """Train the network and periodically reconstruct the test set; returns the filename
    of the latest reconstruction.  Arguments as in the reference (:302-369).

    Differences that do not change results: ``dropout_rate_train`` and ``nprefetch`` are
    accepted and ignored (dropout never fires in the reference; the feed is
    on the GPU); only nearest-neighbour ``resize_method`` exists (anything else raises, see
    INTEGRATION.md); ``tensorboard`` writes the reference's scalars and images (:575-584) through
    ``torch.utils.tensorboard`` when it is importable.

    New capability: inside an initialised ``torch.distributed`` group
    of N > 1 ranks (one process per GPU) training is data parallel -- every rank builds
    ``batch_size`` frames of a global batch of N x ``batch_size`` drawn from one shared index
    stream, gradients are all-reduced (an epoch is ceil(train_len / (N x batch_size)) steps) --
    and the reconstruction sweep is sharded by time range into one output file.  Rank 0 prints,
    saves checkpoints and TensorBoard scalars; every rank returns the same file name.
    """
    import torch
    from .engine import Plan
    from .sampler import TrainSampler, sequential_batches
    from .trainer import Reconstructor, Trainer

    if resize_method not in (NEAREST_NEIGHBOR, "nearest", None):
        raise ValueError("dincae_b200 implements only nearest-neighbour upsampling "
                         "(resize_method=tf.image.ResizeMethod.NEAREST_NEIGHBOR == 1)")
    init_seed, shuffle_seed, noise_seed = 0, None, random.getrandbits(62)
    if iseed is not None:
        # same draw order as :372-375 (np.random -> TF graph seed -> python random)
        np.random.seed(iseed)
        init_seed = int(np.random.randint(0, 2 ** 32 - 1))
        random.seed(int(np.random.randint(0, 2 ** 32 - 1)))
        shuffle_seed = init_seed
        noise_seed = init_seed

    dist, rank, world = _dp_context()
    say = print if rank == 0 else (lambda *a, **k: None)
    py_random = None
    if dist is not None:
        # one shared index / noise / initialisation stream on every rank
        box = [init_seed, shuffle_seed if shuffle_seed is not None else random.getrandbits(31),
               noise_seed, None if iseed is not None else random.getrandbits(62)]
        dist.broadcast_object_list(box, src=0)
        init_seed, shuffle_seed, noise_seed = box[0], box[1], box[2]
        if box[3] is not None:
            py_random = random.Random(box[3])

    say("regularization_L2_beta ", regularization_L2_beta)
    say("enc_nfilter_internal ", enc_nfilter_internal)
    say("nvar ", nvar)
    say("nepoch_keep_missing ", nepoch_keep_missing)

    if outdir is not None and not os.path.isdir(outdir) and rank == 0:
        os.mkdir(outdir)
    if dist is not None:
        dist.barrier()

    jmax, imax = mask.shape
    precision = get_precision()
    plan = Plan(jmax, imax, nvar, enc_nfilter_internal, frac_dense_layer, skipconnections,
                lanes=8 if precision == "bf16" else 4)
    if precision != "bf16":
        from . import _lib
        _lib.check(_lib.load().dincae_set_conv_backend(0 if precision == "fp32" else 1), "set_conv_backend")
    say("precision ", precision)

    if not isinstance(train_datagen, FrameSource) or not isinstance(test_datagen, FrameSource):
        if dist is not None:
            raise ValueError("data-parallel runs need the generators of data_generator / "
                             "data_generator_list (the batch is built on each rank's GPU)")
        from .hostfeed import reconstruct_from_generators
        writer_hf = None
        if tensorboard and outdir is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                writer_hf = SummaryWriter(os.path.join(outdir, "train"))
            except Exception:            # noqa: BLE001
                writer_hf = None
        try:
            return reconstruct_from_generators(
                plan, lon, lat, mask, meandata, train_datagen, train_len, test_datagen, test_len,
                outdir, epochs=epochs, batch_size=batch_size, save_each=save_each,
                truth_uncertain=truth_uncertain, shuffle_buffer_size=shuffle_buffer_size,
                clip_grad=clip_grad, regularization_L2_beta=regularization_L2_beta, transfun=transfun,
                savesample=savesample, learning_rate=learning_rate,
                learning_rate_decay_epoch=learning_rate_decay_epoch, init_seed=init_seed,
                shuffle_seed=shuffle_seed, loss=loss, output_name=_output_name,
                close_writer=_close_writer, iseed=iseed, nepoch_keep_missing=nepoch_keep_missing,
                save_model_each=save_model_each, save_checkpoint=save_checkpoint,
                tensorboard_writer=writer_hf)
        finally:
            if writer_hf is not None:
                writer_hf.close()

    cube = train_datagen.cube()
    test_cube = test_datagen.cube()
    if cube.nvar != nvar:
        raise ValueError("nvar=%d does not match the generator (%d)" % (nvar, cube.nvar))

    trainer = Trainer(plan, cube, batch_size, truth_uncertain=truth_uncertain,
                      clip_grad=clip_grad, l2_beta=regularization_L2_beta,
                      noise_seed=noise_seed, init_seed=init_seed,
                      jitter_std=train_datagen.jitter_std, dist_group=True if dist is not None else None,
                      dropout_rate=dropout_rate_train if _dropout_active[0] else 0.0)
    sampler = TrainSampler(train_len, batch_size * world, shuffle_buffer_size, seed=shuffle_seed,
                           py_random=py_random, train=train_datagen.train)
    mine = slice(rank * batch_size, (rank + 1) * batch_size)
    shard_sweep = dist is not None and not ncio.have_netcdf4()      # HDF5: single writer
    # With the stock `savesample` and no output transform the per-batch arithmetic of
    # savesample (:237-248, 288-291) runs on the device and the sweep hands file-ready records
    # to writer threads; any other hook / transform gets (m_rec, sigma2_rec) on the host like
    # in the reference.
    fast_save = savesample is globals()["savesample"] and transfun[1] is identity
    recon = Reconstructor(plan, test_cube, batch_size, net=trainer.net,
                          train_mode_inputs=test_datagen.train, noise_seed=noise_seed + 1,
                          jitter_std=test_datagen.jitter_std,
                          records=(meandata, -9999.) if fast_save else None,
                          n_out_slots=_SAVE_SLOTS if fast_save else 2)
    writer_tb = None
    if tensorboard and outdir is not None and rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer_tb = SummaryWriter(os.path.join(outdir, "train"))
        except Exception:            # noqa: BLE001
            writer_tb = None

    dt_start = datetime.now()
    say(dt_start)
    fname = None
    index = 0
    steps_per_epoch = ceil(train_len / (batch_size * world))

    # TensorBoard images of the reference (:578-584): m_rec, m_true, sigma2_rec of the first 3
    # frames of every training batch, times the land-sea mask, flipped upside down, scaled to uint8
    # the way tf.summary.image scales float images.  The frames go through a small pinned ring and
    # are written when it is full / at the end of an epoch, so the step loop never synchronises.
    img_k = min(3, batch_size)
    img_ring, img_pending, img_index = None, [], 0
    if writer_tb is not None:
        img_ring = [torch.zeros(3, img_k, jmax, imax).pin_memory() for _ in range(32)]
        issea = torch.from_numpy(np.asarray(mask, dtype=np.float32)).to(trainer.net.dev)

    def snapshot_images():
        nonlocal img_index
        net = trainer.net
        if len(img_pending) == len(img_ring):
            flush_images()
        net.head_recon(img_k)
        xt = net.xtrue[:img_k]
        s2t = 1.0 / torch.clamp(xt[..., 1], min=1e-3)                 # :524-527
        buf = img_ring[len(img_pending)]
        buf[0].copy_(net.m_rec[:img_k] * issea, non_blocking=True)
        buf[1].copy_(xt[..., 0] * s2t * issea, non_blocking=True)
        buf[2].copy_(net.s2_rec[:img_k] * issea, non_blocking=True)
        img_pending.append(img_index)
        img_index += 1

    def flush_images():
        if not img_pending:
            return
        torch.cuda.current_stream().synchronize()
        for slot, step in enumerate(img_pending):
            arr = img_ring[slot].numpy()
            for name, a in zip(("m_rec", "m_true", "sigma2_rec"), arr):
                for j in range(img_k):
                    writer_tb.add_image("Validation/%s/image/%d" % (name, j), _tf_image_u8(a[j][::-1])[None],
                                        step, dataformats="CHW")
        del img_pending[:]

    def drain(e, first_step):
        nonlocal index
        for k, (c, r) in enumerate(trainer.flush_losses()):
            loss.append(c)
            if writer_tb is not None:
                writer_tb.add_scalar("Validation/cost", c, index)
                writer_tb.add_scalar("Validation/RMS", r, index)
            index += 1
            ii = first_step + k
            if ii % 20 == 0:
                say("Epoch: {}/{}...".format(e + 1, epochs),
                    "Training loss: {:.20f}".format(c), "RMS: {:.20f}".format(r), lr_e)

    for e in range(epochs):
        if nepoch_keep_missing > 0:
            random.seed(iseed + e // nepoch_keep_missing)        # :639-641
        lr_e = learning_rate * (0.5 ** (e / learning_rate_decay_epoch))
        trainer.set_lr(lr_e)
        for ii in range(steps_per_epoch):
            fi, ci, seq = sampler.next_batch()
            if world > 1:
                fi, ci, seq = fi[mine], ci[mine], seq[mine]
            trainer.train_step(fi, ci, seq)
            if img_ring is not None:
                snapshot_images()
        drain(e, 0)
        if img_ring is not None:
            flush_images()

        if ((e == epochs - 1) or ((save_each > 0) and (e % save_each == 0))) and outdir is not None:
            say("Save output", e)
            fname = _output_name(outdir)
            if dist is not None:
                box = [fname]
                dist.broadcast_object_list(box, src=0)          # one name (the clock differs by rank)
                fname = box[0]
            if fast_save and (dist is None or shard_sweep):
                if dist is None:
                    _save_records(recon, fname, lon, lat, meandata, test_len, batch_size)
                else:
                    _save_records(recon, fname, lon, lat, meandata, test_len, batch_size, rank, world,
                                  dist.barrier)
            elif fast_save:
                if rank == 0:
                    _save_records(recon, fname, lon, lat, meandata, test_len, batch_size)
                dist.barrier()
            elif rank != 0:
                dist.barrier()                                   # rank 0 runs the hook path alone
            else:
                pending = None
                recon.begin_sweep()
                for ii, frames in enumerate(sequential_batches(test_len, batch_size)):
                    out = recon.run_batch(frames, slot=ii % 2)
                    if pending is not None:
                        _emit(savesample, fname, pending, meandata, lon, lat, e, transfun)
                    ev = torch.cuda.Event()
                    ev.record()
                    pending = (out, len(frames), ii, ii * batch_size, ev)
                if pending is not None:
                    _emit(savesample, fname, pending, meandata, lon, lat, e, transfun)
                _close_writer(fname)
                if dist is not None:
                    dist.barrier()

        if (save_model_each > 0) and (e % save_model_each == 0) and outdir is not None and rank == 0:
            save_checkpoint(os.path.join(outdir, "model-{:03d}.ckpt".format(e + 1)), trainer)

    if writer_tb is not None:
        writer_tb.close()
    dt_end = datetime.now()
    say(dt_end)
    say(dt_end - dt_start)
    return fname


def _tf_image_u8(img):
    """float image -> uint8 like tf.summary.image (tensorflow/core/kernels/summary_image_op.cc):
    all values >= 0: scaled so that the largest is 255; otherwise 0 maps to 128 and the largest
    magnitude to +-127."""
    img = np.asarray(img, dtype=np.float32)
    lo, hi = float(img.min()), float(img.max())
    if lo < 0:
        mx = max(abs(lo), abs(hi))
        scale, offset = (127.0 / mx if mx > 1e-6 else 0.0), 128.0
    else:
        scale, offset = (255.0 / hi if hi > 1e-6 else 0.0), 0.0
    return np.clip(img * scale + offset, 0, 255).astype(np.uint8)


_SAVE_SLOTS, _SAVE_THREADS = 6, 3


def shard_batches(test_len, batch_size, rank=0, world=1):
    """Batch numbers [b0, b1) of the reconstruction sweep that ``rank`` takes: contiguous time
    ranges, whole batches, no overlap (no collective on this path)."""
    nb = ceil(test_len / batch_size)
    per = ceil(nb / world)
    return min(nb, rank * per), min(nb, (rank + 1) * per)


def _save_records(recon, fname, lon, lat, meandata, test_len, batch_size, rank=0, world=1,
                  barrier=None):
    """Reconstruction sweep of the stock output path (:671-689): the device emits the file's
    records batch by batch into a ring of pinned buffers; writer threads wait for a batch's
    copy event and put it into the file with one positional write.  The file is complete and
    closed on return.

    ``world > 1`` (one process per GPU): every rank reconstructs its own contiguous time range
    and writes it into the SAME file at the records' own offsets -- rank 0 lays down the header
    and the fixed variables first and patches the record count last; ``barrier()`` separates
    the three phases.  No data moves between ranks."""
    from concurrent.futures import ThreadPoolExecutor
    sharded = world > 1
    if sharded and barrier is None:
        raise ValueError("a sharded sweep needs a barrier")
    writer = None
    if rank == 0:
        writer = ncio.ReconWriter(fname, lon, lat, meandata, expect_records=test_len)
    if sharded:
        barrier()                                      # the header exists
        if rank != 0:
            writer = ncio.ReconWriter(fname, lon, lat, meandata, append=True)
    nslots = len(recon.out_host)
    b0, b1 = shard_batches(test_len, batch_size, rank, world)

    def put(ev, buf, offset, n):
        ev.synchronize()
        writer.write_records(offset, buf.numpy().reshape(-1).view(np.uint8), n)

    futures = [None] * nslots
    if hasattr(recon, "begin_sweep"):
        recon.begin_sweep()
    if getattr(recon, "use_graph", False):
        # graphs of every batch size of this range (full batches + the tail) are captured before
        # the writer threads start: a capture must not overlap their event waits
        for n in sorted({min(test_len, (ii + 1) * batch_size) - ii * batch_size for ii in range(b0, b1)}):
            recon.capture(n)
    try:
        with ThreadPoolExecutor(max_workers=_SAVE_THREADS) as pool:
            for k, ii in enumerate(range(b0, b1)):
                frames = np.arange(ii * batch_size, min(test_len, (ii + 1) * batch_size), dtype=np.int32)
                slot = k % nslots
                if futures[slot] is not None:
                    futures[slot].result()             # the slot's previous batch is in the file
                out = recon.run_batch(frames, slot=slot)
                futures[slot] = pool.submit(put, recon.record_event(), out, ii * batch_size, len(frames))
            for f in futures:
                if f is not None:
                    f.result()
    finally:
        if sharded:
            if rank != 0:
                writer.close(finalize=False)
            barrier()                                  # every rank's records are in the file
            if rank == 0:
                writer.numrecs = test_len
                writer.close()
        else:
            writer.close()


def _emit(savesample_fn, fname, pending, meandata, lon, lat, e, transfun):
    out, n, ii, offset, ev = pending
    ev.synchronize()
    arr = out.numpy()
    savesample_fn(fname, arr[0, :n], arr[1, :n], meandata, lon, lat, e, ii, offset,
                  transfun=transfun)


def save_checkpoint(path, trainer):
    """Counterpart of ``saver.save`` (:691-693): the 20 variables in TF order/shape plus the
    Adam slots, as a torch file ``<path>.pt`` (the reference has no restore path)."""
    import torch
    net = trainer.net
    torch.save({
        "names": [e["name"] for e in net.plan.entries],
        "params": net.get_params_tf(),
        "adam_m": net.plan.unpack(net.adam_m.cpu().numpy()),
        "adam_v": net.plan.unpack(net.adam_v.cpu().numpy()),
        "adam_state": net.adam_state.cpu().numpy(),
        "step": trainer.step_count,
    }, path + ".pt")
    return path + ".pt"


def load_checkpoint(path, trainer):
    """Restore what ``save_checkpoint`` wrote (variables, Adam slots and running beta powers,
    step counter) into a trainer built for the same architecture -- the resume path the
    reference lacks.  ``path`` with or without the ``.pt`` suffix."""
    import numpy as np
    import torch
    if not path.endswith(".pt"):
        path = path + ".pt"
    ck = torch.load(path, map_location="cpu", weights_only=False)
    net = trainer.net
    names = [e["name"] for e in net.plan.entries]
    if list(ck["names"]) != names:
        raise ValueError("checkpoint variables %s do not match the network %s" % (ck["names"][:3], names[:3]))
    net.set_params_tf(ck["params"])
    net.adam_m.copy_(torch.from_numpy(net.plan.pack(ck["adam_m"])))
    net.adam_v.copy_(torch.from_numpy(net.plan.pack(ck["adam_v"])))
    net.adam_state.copy_(torch.from_numpy(np.asarray(ck["adam_state"], dtype=np.float32)))
    trainer.step_count = int(ck["step"])
    return trainer


def reconstruct_gridded_nc(filename, varname, outdir, jitter_std=0.05, ntime_win=3,
                           transfun=(identity, identity), **kwargs):
    """(:704-740).  Bug-compatible with the reference: ``transfun[0]`` is not applied to the
    input (Q5), the "test" generator is built in train mode (Q1), and ``None`` is returned
    (Q7)."""
    lon, lat, time, data, missing, mask = load_gridded_nc(filename, varname)
    train_datagen, nvar, train_len, meandata = data_generator(
        lon, lat, time, data, missing, ntime_win=ntime_win, jitter_std=jitter_std)
    test_datagen, nvar, test_len, test_meandata = data_generator(
        lon, lat, time, data, missing, ntime_win=ntime_win, train=False)
    print("Number of input variables: ", nvar)
    reconstruct(lon, lat, mask, meandata, train_datagen, train_len, test_datagen, test_len,
                outdir, transfun=transfun, nvar=nvar, **kwargs)


def reconstruct_gridded_files(fields, outdir, ntime_win=3, **kwargs):
    """(:744-811) multivariate driver; reconstructs the first field, returns the filename."""
    n = len(fields)
    data_full, missing = [None] * n, [None] * n
    jitter_std, transfun = [None] * n, [None] * n
    obs_err_std = [1] * n
    for i, field in enumerate(fields):
        transfun[i] = field.get("transfun", (identity, identity))
        (field["lon"], field["lat"], field["time"], field["data"], field["missing"],
         field["mask"]) = load_gridded_nc(field["filename"], field["varname"])
        data_full[i] = transfun[i][0](field["data"])
        missing[i] = field["missing"]
        jitter_std[i] = field.get("jitter_std", 0)
    lon, lat = fields[0]["lon"], fields[0]["lat"]
    time, mask = fields[0]["time"], fields[0]["mask"]
    train_datagen, nvar, train_len, meandata = data_generator_list(
        lon, lat, time, data_full, missing, obs_err_std=obs_err_std, jitter_std=jitter_std,
        ntime_win=ntime_win)
    test_datagen, nvar, test_len, test_meandata = data_generator_list(
        lon, lat, time, data_full, missing, obs_err_std=obs_err_std, ntime_win=ntime_win,
        train=False)
    print("Number of input variables: ", nvar)
    return reconstruct(lon, lat, mask, meandata, train_datagen, train_len, test_datagen,
                       test_len, outdir, transfun=transfun[0], nvar=nvar, **kwargs)
