"""Latent sampling (reference graph_util.py:5-18): the numpy ``RandomState`` stream is the
bit-exact input definition of the path."""
import numpy as np

from . import constants


def z_sample(batch_size, seed=0, dim_z=constants.DIM_Z):
    return np.random.RandomState(seed).randn(batch_size, dim_z)


def w_sample(batch_size, seed=0, dim_z=constants.DIM_Z):
    return np.random.RandomState(seed).uniform(low=-1, high=2, size=(batch_size, dim_z))


def graph_input(graph, num_samples, seed=0, **kwargs):
    return {"z": z_sample(num_samples, seed, graph.dim_z)}
