"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's GAN2Shape/renderer path (torch-CPU elementwise ops) and of the
external neural_renderer rasteriser (plain C).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything from here; the product package
(gan-2d-to-3d_b200/) never does.
"""
