"""CPU oracle for the VAE-GAN training step -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain fp32 PyTorch on the CPU, the algorithm of the
reference's hot path (Andrey1408/vae-gan-mark: the conv encoder / decoder /
U-Net+FiLM generator, the PatchGAN discriminator, the KL / L1 / hinge losses and
the per-batch step body).  It exists so that the CUDA path in
``vae_gan_mark_b200`` can be checked against something that travels to the GPU
box (``/root/reference`` does not).

Rules ("Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import anything from here;
  * the product package never imports it and has no CPU fallback;
  * every function cites the reference file:line it follows;
  * the oracle is PINNED: ``tests/golden/make_golden.py`` imports the reference's own
    ``nn.Module`` classes from ``/root/reference`` (in the build container) and records
    their outputs; ``tests/test_oracle_golden.py`` checks this restatement against
    those fixtures bit-for-tolerance (fp32, rtol 1e-5).

Unpinned pieces (cannot be obtained offline, so they are stubbed identically on
both sides): the SBERT sentence embedder of ``vae-gan.py:86-116`` (a seeded hash
embedding stands in) and the ImageNet VGG16 perceptual term
(``vae-gan.py:300-311``; weight 0 in every comparison).
"""
