"""Drop-in surface of the reference's ``models/pggan/pggan.py`` (class PGGAN, :12-386) on the B200-native engine
(SURVEY.md 8f, row f4).  Same constructor ``PGGAN(batch_size, steps, check_dir_write, check_dir_read, dataset,
sample_path, log_dir, stage, trans, build_model=True)``, the same attribute names for the inputs (``iter,
learning_rate, x, x_mismatch, cond, z, epsilon, z_sample, cond_sample``), the fetchable tensors / scalars / train ops
(``G, mean, log_sigma, Dg_logit, ..., sampler, D_loss, G_loss, D_optim, G_optim``), ``d_vars / g_vars``, the stage
helpers (``get_nf, get_dnf, get_rgb_name, get_conv_scope_name, get_variables_up_to_stage``) and ``train()`` with the
reference's step order, sampling / checkpoint periods and the stage-to-stage restore (:147-250).
``run(fetches, feed_dict)`` stands in for ``sess.run``.

All arithmetic runs in the CUDA library through ``t2i_b200.kernels`` (engine_pggan.PgganEngine); there is no CPU
fallback.  Summaries / PNG writing depend on TF and removed scipy APIs and are out of scope; the loss scalars printed
every 20 iterations (:221-226) go to ``self.log``.
"""
import os
import time

import numpy as np
import torch

from ...engine_pggan import ADAM_LR, PgganEngine
from ...utils.saver import load, save
from ..wgancls.model import Fetch, Placeholder, _truncated_normal


class _VarSubset(object):
    """tf.train.Saver(var_list) (pggan.py:122-134): utils/saver.py's helpers see the variables (and, as
    tf.global_variables(scope) does, their Adam slots) whose names start with one of ``prefixes``."""

    def __init__(self, model, prefixes):
        self.model, self.prefixes = model, tuple(prefixes)

    def get_variables(self):
        return {k: v for k, v in self.model.get_variables().items() if k.startswith(self.prefixes)}

    def set_variables(self, variables):
        self.model.set_variables({k: v for k, v in variables.items() if k.startswith(self.prefixes)})

    def get_optimizer_state(self):
        st = self.model.get_optimizer_state()
        return {k: v for k, v in st.items() if k[2:].startswith(self.prefixes)}      # 'm/<var>', 'v/<var>'

    def set_optimizer_state(self, state):
        self.model.set_optimizer_state({k: v for k, v in state.items() if k[2:].startswith(self.prefixes)})


class PGGAN(object):

    # build model
    def __init__(self, batch_size, steps, check_dir_write, check_dir_read, dataset, sample_path, log_dir, stage, trans,
                 build_model=True, precision="bf16", device=None, kernels=None, distributed=None, use_graphs=True,
                 nf_base=1024, nf_cap=512, z_dim=128, embed_dim=1024, compr_embed_dim=128, sample_num=64, d_embed=128):
        """The first ten arguments are the reference's (pggan.py:15-16).  precision / device / kernels / distributed /
        use_graphs as for WGanCls; the remaining keywords default to the reference's literals (:28-38, :339-343)."""
        self.batch_size = batch_size
        self.steps = steps
        self.check_dir_write = check_dir_write
        self.check_dir_read = check_dir_read
        self.dataset = dataset
        self.sample_path = sample_path
        self.log_dir = log_dir
        self.stage = stage
        self.trans = trans

        self.z_dim = z_dim
        self.embed_dim = embed_dim
        self.out_size = 4 * pow(2, stage - 1)
        self.channel = 3
        self.sample_num = sample_num
        self.compr_embed_dim = compr_embed_dim
        self.lr = 0.00005
        self.lr_inp = self.lr
        self.output_size = 4 * pow(2, stage - 1)

        self.alpha_tra = 0.0
        self._nf_base, self._nf_cap, self._d_embed = nf_base, nf_cap, d_embed      # d_embed: the literal 128 of pggan.py:271
        self.precision = precision
        self._np = {"bf16": 1, "bf16x3": 2}[precision]
        if kernels is None:
            from ... import kernels as _k      # loads libt2i_b200.so lazily, raises if missing
            kernels = _k
        self._K = kernels
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("PGGAN needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self._world, self._allreduce = 1, None
        if distributed:
            import torch.distributed as dist
            group = None if distributed is True else distributed
            self._world = dist.get_world_size(group)
            self._allreduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        self._engines = {}
        self._use_graphs = use_graphs
        self._noise_gen = None
        self._built = False
        self.log = []
        self._train_engine()
        if build_model:
            self.build_model()
            self.define_losses()

    def _engine(self, batch):
        if batch not in self._engines:
            base = next(iter(self._engines.values()), None)
            self._engines[batch] = PgganEngine(
                self._K, self.device, batch, self._np, self.stage, self.trans, self.z_dim, self.embed_dim,
                self.compr_embed_dim, self._nf_base, self._nf_cap, self._d_embed, 9, self._world, self._allreduce,
                share_from=base, use_graphs=self._use_graphs)
        return self._engines[batch]

    def _train_engine(self):
        return self._engine(self.batch_size)

    def build_model(self):
        """pggan.py:49-84: the placeholders and the fetchable tensors of the reference graph, by name."""
        b, s = self.batch_size, self.output_size
        self.iter = Placeholder("iter")
        self.learning_rate = Placeholder("learning_rate")     # fed and unused, as in the reference (:111-112)
        self.x = Placeholder("x", [b, s, s, self.channel])
        self.x_mismatch = Placeholder("x_mismatch", [b, s, s, self.channel])
        self.cond = Placeholder("cond", [b, self.embed_dim])
        self.z = Placeholder("z", [b, self.z_dim])
        # the reference rebinds self.epsilon to a tf.random_uniform tensor (:68) and the trainer FEEDS that tensor
        # (:205): the fed value is what the graph uses.  Not fed -> drawn here.
        self.epsilon = Placeholder("eps", [b, 1, 1, 1])
        self.z_sample = Placeholder("z_sample", [self.sample_num, self.z_dim])
        self.cond_sample = Placeholder("cond_sample", [self.sample_num, self.embed_dim])
        # explicit noise inputs (drawn inside the graph in the reference, :351); optional
        self.cond_noise = Placeholder("cond_noise", [b, self.compr_embed_dim])
        self.cond_noise_sample = Placeholder("cond_noise_sample", [self.sample_num, self.compr_embed_dim])
        for n in ("G", "mean", "log_sigma", "Dg_logit", "Dx_logit", "Dxmi_logit", "x_hat", "Dx_hat_logit", "sampler"):
            setattr(self, n, Fetch(n, "tensor"))
        self.cond_inp = self.cond
        self.alpha_assign = Fetch("alpha_assign", "op")
        self.d_vars = [n for n in self.variable_names() if n.startswith("d_net/")]
        self.g_vars = [n for n in self.variable_names() if n.startswith("g_net/")]
        self._built = True

    def define_losses(self):
        """pggan.py:94-134: loss scalars, the two train ops and the two savers."""
        for n in ("D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss", "G_kl_loss", "real_gp",
                  "real_gp2", "D_loss", "G_loss"):
            setattr(self, n, Fetch(n, "scalar"))
        self.D_optim = Fetch("D_optim", "op")
        self.G_optim = Fetch("G_optim", "op")
        self.saver = _VarSubset(self, self.get_variables_up_to_stage(self.stage))                 # :122-126
        self.restore = None
        if self.stage > 1 and self.trans:                                                        # :129-134
            self.restore = _VarSubset(self, self.get_variables_up_to_stage(self.stage - 1))

    # ------------------------------------------------------------------ stage helpers (pggan.py:330-345,380-386)
    def get_rgb_name(self, stage):
        return 'rgb_stage_%d' % stage

    def get_conv_scope_name(self, stage):
        return 'conv_stage_%d' % stage

    def get_dnf(self, stage):
        return min(self._nf_base // (2 ** stage) * 2, self._nf_cap)

    def get_nf(self, stage):
        return min(self._nf_base // (2 ** stage) * 4, self._nf_cap)

    def get_variables_up_to_stage(self, stages):
        """Scope prefixes of the variables the reference collects (:380-386)."""
        pre = ['d_net/%s/' % self.get_rgb_name(stages - 1), 'g_net/%s/' % self.get_rgb_name(stages - 1)]
        for stage in range(stages):
            pre += ['d_net/%s/' % self.get_conv_scope_name(stage), 'g_net/%s/' % self.get_conv_scope_name(stage)]
        return pre

    # ------------------------------------------------------------------ variables (checkpoint boundary)
    def variable_names(self):
        return list(self._train_engine().get_params_tf().keys())

    def get_variables(self):
        return self._train_engine().get_params_tf()

    def set_variables(self, variables):
        """A partial dict is merged into the current values (restoring the previous stage leaves this stage's new
        layers at their initial values, pggan.py:147-160)."""
        eng = self._train_engine()
        cur = eng.get_params_tf()
        cur.update({k: torch.as_tensor(np.asarray(v)) for k, v in variables.items() if k in cur})
        eng.set_params_tf(cur)

    def get_optimizer_state(self):
        return self._train_engine().get_adam_tf()

    def set_optimizer_state(self, state):
        """Partial Adam slots are merged; the step counters are NOT restored across stages: the beta-power
        accumulators live outside the saved scopes (:380-386) and restart with every new graph."""
        if not state:
            return
        eng = self._train_engine()
        cur = eng.get_adam_tf()
        cur.update({k: torch.as_tensor(np.asarray(v)) for k, v in state.items() if k in cur and k not in ("d_t", "g_t")})
        cur["d_t"], cur["g_t"] = eng.d_t, eng.g_t
        eng.set_adam_tf(cur)

    def initialize(self, seed=0):
        """initialize_uninitialized (pggan.py:163-164) for a fresh graph: He truncated-normal kernels
        (utils/ops.py:60,86), zero biases / beta, gamma 1."""
        gen = torch.Generator().manual_seed(seed)
        eng = self._train_engine()
        p = eng.get_params_tf()
        for name, w in p.items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("weights", "kernel"):
                fan_in = w.shape[-2] * int(np.prod(w.shape[:-2])) if w.dim() > 1 else 1
                t = torch.empty(w.shape)
                torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=gen)
                p[name] = t * (1.3 * 2.0 / fan_in) ** 0.5
            elif leaf == "gamma":
                p[name] = torch.ones_like(w)
            else:
                p[name] = torch.zeros_like(w)
        eng.set_params_tf(p)
        for k in ("d", "g"):
            eng.adam_m[k].zero_()
            eng.adam_v[k].zero_()
        eng.d_t = eng.g_t = 0

    # ------------------------------------------------------------------ eager sub-graphs
    def _dev(self, a, shape=None):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
        t = t.to(self.device, torch.float32, non_blocking=True)
        return t.reshape(shape) if shape is not None else t

    def seed_noise(self, seed):
        self._noise_gen = torch.Generator(device=self.device).manual_seed(seed)

    def _check_graph(self, stages, t):
        if stages != self.stage or bool(t) != bool(self.trans):
            raise NotImplementedError("this model object holds the graph of stage %d (trans=%s) only" % (self.stage, self.trans))

    def generator(self, z_var, cond_inp, stages=None, t=None, reuse=False, cond_noise=True, noise=None, alpha=None):
        """pggan.py:279-316.  Returns (image NHWC [B, S, S, 3], mean, log_sigma) as CUDA fp32 tensors."""
        self._check_graph(self.stage if stages is None else stages, self.trans if t is None else t)
        z = self._dev(z_var)
        b = z.shape[0]
        eng = self._engine(b)
        eng.set_alpha(self.alpha_tra if alpha is None else alpha)
        cond = self._dev(cond_inp, (b, self.embed_dim))
        tn = self._dev(noise) if noise is not None else _truncated_normal((b, self.compr_embed_dim), self.device,
                                                                         self._noise_gen)
        s = self.output_size
        out = torch.empty(b, s, s, 3, device=self.device, dtype=torch.float32)
        eng.g["kl_scratch"].zero_()
        eng.g_forward(z, cond, tn, out, eng.g["kl_scratch"], cond_noise=cond_noise)
        ce = self.compr_embed_dim
        ms = eng.g["ms"].clone()        # fp32 [b, 2*ce] = [mean | log_sigma]
        return out, ms[:, :ce], ms[:, ce:]

    def discriminator(self, inp, cond, stages=None, t=None, reuse=False, alpha=None):
        """pggan.py:251-277.  inp NHWC [B, S, S, 3], cond [B, 1024] -> logits [B, 1, 1, 1]."""
        self._check_graph(self.stage if stages is None else stages, self.trans if t is None else t)
        x = self._dev(inp)
        b = x.shape[0]
        eng = self._engine(b)
        eng.set_alpha(self.alpha_tra if alpha is None else alpha)
        eng.d["img"][:b].copy_(x)
        self._K.to_planes(self._dev(cond, (b, self.embed_dim)), eng.d["cond"][:, :b])
        eng.d_forward(0, b)
        return eng.d["logit"][:b].clone().reshape(b, 1, 1, 1)

    # ------------------------------------------------------------------ sess.run stand-in
    def run(self, fetches, feed_dict=None):
        """``sess.run(fetches, feed_dict)`` for the fetch lists of train(): [D_optim, D_loss] (:218),
        [G_optim, G_loss] (:219), sampler (:230) and the summary scalars (:137-158)."""
        if not self._built:
            raise RuntimeError("the model was constructed with build_model=False")
        feed_dict = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        names = [f.name for f in flist]
        eng = self._train_engine()
        b = self.batch_size
        get = lambda k: feed_dict[k] if k in feed_dict else None
        t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a)

        def stage_feed(need_images):
            noise = get(self.cond_noise)
            if noise is None:       # a fresh truncated-normal draw per run (:351)
                noise = _truncated_normal((b, self.compr_embed_dim), self.device, self._noise_gen)
            eps = get(self.epsilon) if need_images else None
            if need_images and eps is None:
                eps = torch.rand(b, device=self.device, generator=self._noise_gen)      # :68
            eng.load_feed(x=t(get(self.x)) if need_images else None,
                          x_mismatch=t(get(self.x_mismatch)) if need_images else None, cond=t(get(self.cond)),
                          z=t(get(self.z)), epsilon=t(eps), tn_eps=t(noise))

        ran = False
        if "D_optim" in names or "alpha_assign" in names:
            it = get(self.iter)
            if it is not None:                                          # alpha_assign, :78-79,118-119
                self.alpha_tra = float(it) / float(self.steps)
        if "D_optim" in names:
            stage_feed(True)
            eng.d_step(self.alpha_tra, ADAM_LR)
            ran = True
        if "G_optim" in names:
            stage_feed(False)
            eng.set_alpha(self.alpha_tra)
            eng.g_step(ADAM_LR)
            ran = True
        out, sc = [], None
        for f in flist:
            if f.kind == "op":
                out.append(None)
            elif f.kind == "scalar":
                if not ran and sc is None:
                    raise RuntimeError("scalar '%s' is produced by the D/G run; fetch it with the train op" % f.name)
                sc = sc or eng.scalars_dict()
                out.append(sc[f.name])
            elif f.name == "sampler":
                # The generator normalises per sample (no batch statistics): the sample_num images are produced in chunks
                # of the training batch on the TRAINING engine -- a second engine for sample_num images would allocate
                # its own discriminator and gradient buffers (tens of GB at 256 / 512 pixels) and never use them.
                zs = np.asarray(feed_dict[self.z_sample], dtype=np.float32)
                cs = np.asarray(feed_dict[self.cond_sample], dtype=np.float32).reshape(zs.shape[0], -1)
                ns = get(self.cond_noise_sample)
                ns = None if ns is None else np.asarray(ns, dtype=np.float32)
                n, imgs = zs.shape[0], []
                for i in range(0, n, b):
                    idx = np.arange(i, i + b) % n                       # the last chunk wraps around to fill the batch
                    img, _, _ = self.generator(zs[idx], cs[idx], noise=None if ns is None else ns[idx])
                    imgs.append(img[:min(b, n - i)].cpu().numpy())
                out.append(np.concatenate(imgs, 0))
            elif f.name == "G":
                out.append(eng.d["img"][:b].cpu().numpy())
            elif f.name == "x_hat":
                out.append(eng.d["img"][3 * b:].cpu().numpy())
            elif f.name in ("Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"):
                k = ["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"].index(f.name)
                out.append(eng.d["logit"][k * b:(k + 1) * b].cpu().numpy().reshape(b, 1, 1, 1))
            else:
                raise KeyError("fetch '%s' is not materialised by this implementation" % f.name)
        return out[0] if single else out

    # do train
    def train(self, max_updates=None, on_samples=None):
        """pggan.py:160-250: restore (previous stage while `trans`, else this stage), then for idx in 1 .. steps-1 the
        D run and the G run on the same feed; losses every 20, samples and a checkpoint every 2000 and at the end."""
        start_point = 0
        self.initialize()
        if self.stage != 1:
            if self.trans:
                could_load, _ = load(self.restore, self.check_dir_read, prefix="pggan")
                if not could_load:
                    raise RuntimeError('Could not load previous stage during transition')
            else:
                could_load, _ = load(self.saver, self.check_dir_read, prefix="pggan")
                if not could_load:
                    raise RuntimeError('Could not load current stage')

        sample_z = np.random.normal(0, 1, (self.sample_num, self.z_dim))
        _, sample_cond, _, captions = self.dataset.test.next_batch_test(self.sample_num, 0, 1)
        sample_cond = np.squeeze(sample_cond, axis=0)
        print('Conditionals sampler shape: {}'.format(sample_cond.shape))
        start_time = time.time()
        done = 0
        for idx in range(start_point + 1, self.steps):
            if self.trans:
                self.lr_inp = self.lr
            epoch_size = self.dataset.train.num_examples // self.batch_size
            epoch = idx // epoch_size

            images, wrong_images, embed, _, _ = self.dataset.train.next_batch(self.batch_size, 4, wrong_img=True,
                                                                              embeddings=True)
            batch_z = np.random.normal(0, 1, (self.batch_size, self.z_dim))
            eps = np.random.uniform(0., 1., size=(self.batch_size, 1, 1, 1))

            feed_dict = {
                self.x: images,
                self.learning_rate: self.lr_inp,
                self.x_mismatch: wrong_images,
                self.cond: embed,
                self.z: batch_z,
                self.epsilon: eps,
                self.z_sample: sample_z,
                self.cond_sample: sample_cond,
                self.iter: idx,
            }

            _, err_d = self.run([self.D_optim, self.D_loss], feed_dict=feed_dict)
            _, err_g = self.run([self.G_optim, self.G_loss], feed_dict=feed_dict)

            if np.mod(idx, 20) == 0:
                self.log.append({"epoch": epoch, "idx": idx, "time": time.time() - start_time, "d_loss": err_d,
                                 "g_loss": err_g, "alpha": self.alpha_tra})
                print("Epoch: [%2d] [%4d] time: %4.4f, d_loss: %.8f, g_loss: %.8f"
                      % (epoch, idx, time.time() - start_time, err_d, err_g))

            if np.mod(idx, 2000) == 0:
                try:                                  # pggan.py:228-244: a failed sample is logged, the pass goes on
                    samples = self.run(self.sampler, feed_dict={self.z_sample: sample_z, self.cond_sample: sample_cond})
                    samples = np.clip(samples, -1., 1.)
                    if self.out_size > 256:
                        samples = samples[:4]
                    if on_samples is not None:
                        on_samples(epoch, idx, samples, captions)
                except Exception as e:
                    print("Failed to generate sample image")
                    print(type(e))
                    print(e.args)
                    print(e)

            done += 1
            last = idx == self.steps - 1 or (max_updates is not None and done >= max_updates)
            if np.mod(idx, 2000) == 0 or last:
                save(self.saver, self.check_dir_write, idx, 2, prefix="pggan")
            if last:
                break
