"""The caller of the hot path: mirror of the reference's ``models/wgancls/trainer.py``
(WGanClsTrainer, :12-127) on the B200 model.  Same constructor signature (``sess`` is accepted and
ignored), same step order (D + kt run, then the G run when idx % N_CRITIC == 0, :97-102), same lr
schedule (0.95 ** ((idx // n_critic) // 10000), :82-86), same periods for summaries / samples /
checkpoints (:104-126).  Summary scalars come from the runs themselves (no extra forward as at :106);
sample grids are returned to an optional callback instead of being written as PNGs (the reference's
image utilities depend on removed scipy/imageio APIs and are out of scope)."""
import sys
import time

import numpy as np

from ...utils.saver import load, save


class SyntheticTextDataset(object):
    """Stand-in for preprocess/dataset.py's TextDataset (:263-285) with the two methods the trainer
    calls (``train.next_batch``, ``test.next_batch_test``), producing the synthetic batches of the
    benchmark: images ~ U(-1, 1), 1024-d embeddings ~ N(0, 1)."""

    class _Split(object):
        def __init__(self, num_examples, embed_dim, seed, size=64):
            self.num_examples = num_examples
            self._size = size
            self._embed_dim = embed_dim
            self._rng = np.random.RandomState(seed)

        def next_batch(self, batch_size, window=None, wrong_img=False, embeddings=False, labels=False):
            img = self._rng.uniform(-1, 1, (batch_size, self._size, self._size, 3)).astype(np.float32)
            wrong = self._rng.uniform(-1, 1, (batch_size, self._size, self._size, 3)).astype(np.float32) if wrong_img else None
            emb = self._rng.normal(0, 1, (batch_size, self._embed_dim)).astype(np.float32) if embeddings else None
            return [img, wrong, emb, None, None]

        def next_batch_test(self, batch_size, start, max_captions):
            img = self._rng.uniform(-1, 1, (batch_size, self._size, self._size, 3)).astype(np.float32)
            emb = self._rng.normal(0, 1, (1, batch_size, self._embed_dim)).astype(np.float32)
            return img, emb, None, [["synthetic caption %d" % i] for i in range(batch_size)]

    def __init__(self, embed_dim=1024, num_examples=8192, seed=0, image_size=64, rank=None):
        """rank: data-parallel replicas must draw DIFFERENT batches (each rank is a shard of the global batch): the
        training stream is seeded per rank (default: the torch.distributed rank when a process group exists); the test
        split, which conditions the sample grids, is the same everywhere."""
        if rank is None:
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()
            except Exception:
                pass
        self.train = self._Split(num_examples, embed_dim, seed + 1000 * rank, image_size)
        self.test = self._Split(num_examples // 8, embed_dim, seed + 1, image_size)


class WGanClsTrainer(object):
    def __init__(self, sess, model, dataset, cfg, log=None, on_samples=None):
        self.sess = sess            # ignored: there is no session, the model runs itself
        self.model = model
        self.dataset = dataset
        self.cfg = cfg
        self.lr_d = self.cfg.TRAIN.D_LR
        self.lr_g = self.cfg.TRAIN.G_LR
        self.log = log if log is not None else []
        self.on_samples = on_samples

    def define_summaries(self):
        """trainer.py:20-47: the scalars written every SUMMARY_PERIOD iterations."""
        m = self.model
        self.summary_fetches = [m.D_loss_real, m.D_loss_fake, m.real_gp, m.D_loss, m.reg_loss, m.wdist, m.wdist2,
                                m.D_loss_mismatch, m.real_gp2, m.kt, m.balance_loss]

    def train(self, max_steps=None):
        self.define_summaries()
        m, cfg = self.model, self.cfg
        sample_z = np.random.normal(0, 1, (m.sample_num, m.z_dim))
        _, sample_cond, _, captions = self.dataset.test.next_batch_test(m.sample_num, 0, 1)
        sample_cond = np.squeeze(sample_cond, axis=0)
        print('Conditionals sampler shape: {}'.format(sample_cond.shape))

        start_time = time.time()
        m.initialize()
        could_load, checkpoint_counter = load(m, cfg.CHECKPOINT_DIR)
        start_point = checkpoint_counter if could_load else 0
        print(" [*] Load SUCCESS" if could_load else " [!] Load failed...")
        sys.stdout.flush()

        max_steps = cfg.TRAIN.MAX_STEPS if max_steps is None else max_steps
        for idx in range(start_point + 1, max_steps):
            images, wrong_images, embed, _, _ = self.dataset.train.next_batch(m.batch_size, 4, embeddings=True,
                                                                              wrong_img=True)
            batch_z = np.random.normal(0, 1, (m.batch_size, m.z_dim))
            eps = np.random.uniform(0., 1., size=(m.batch_size, 1, 1, 1))
            n_critic = cfg.TRAIN.N_CRITIC
            kiter = (idx // n_critic) // 10000

            feed_dict = {
                m.learning_rate_d: self.lr_d * (0.95 ** kiter),
                m.learning_rate_g: self.lr_g * (0.95 ** kiter),
                m.x: images,
                m.x_mismatch: wrong_images,
                m.cond: embed,
                m.z: batch_z,
                m.epsilon: eps,
                m.z_sample: sample_z,
                m.cond_sample: sample_cond,
                m.iter: idx,
            }
            summary = np.mod(idx, cfg.TRAIN.SUMMARY_PERIOD) == 0
            fetched = m.run([m.D_optim, m.kt_optim, m.D_loss] + (self.summary_fetches if summary else []),
                            feed_dict=feed_dict)
            err_d = fetched[2]
            err_g = None
            if idx % n_critic == 0:
                _, err_g, kl = m.run([m.G_optim, m.G_loss, m.G_kl_loss], feed_dict=feed_dict)
            if summary:
                rec = {"idx": idx, "D_loss": err_d, "G_loss": err_g, "time": time.time() - start_time}
                rec.update({f.name: v for f, v in zip(self.summary_fetches, fetched[3:])})
                self.log.append(rec)

            if np.mod(idx, cfg.TRAIN.SAMPLE_PERIOD) == 0:
                try:
                    samples = m.run(m.sampler, feed_dict={m.z_sample: sample_z, m.cond_sample: sample_cond})
                    if self.on_samples is not None:
                        self.on_samples(idx, samples, captions)
                except Exception as e:
                    print("Failed to generate sample image")
                    print(type(e))
                    print(e.args)
                    print(e)

            if np.mod(idx, 500) == 2:       # rank 0 writes, every rank passes the barrier inside save()
                save(m, cfg.CHECKPOINT_DIR, idx, cfg.TRAIN.CHECKPOINTS_TO_KEEP)
            sys.stdout.flush()
