"""Nodes from files - input adaptors (/root/reference/src/anemoi/graphs/nodes/builders/from_file.py).

Pure I/O; kept so unchanged recipes can feed the GPU edge path.  ``ZarrDatasetNodes`` imports ``anemoi.datasets``
when it is used (the package is not in this image; the reference imports it at module load)."""

from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch

from ...generate.masks import KNNAreaMaskBuilder
from .base import BaseNodeBuilder

LOGGER = logging.getLogger(__name__)


def open_dataset(*args, **kwargs):
    """``anemoi.datasets.open_dataset``, imported on first use."""
    try:
        from anemoi.datasets import open_dataset as _open
    except ImportError as err:  # pragma: no cover - depends on the environment
        raise ImportError(
            "ZarrDatasetNodes needs the anemoi-datasets package (`pip install anemoi-datasets`), as in the reference."
        ) from err
    return _open(*args, **kwargs)


class ZarrDatasetNodes(BaseNodeBuilder):
    """Nodes from Zarr dataset (from_file.py:28-63).

    Attributes
    ----------
    dataset : str | DictConfig
        The dataset.
    """

    def __init__(self, dataset, name: str) -> None:
        LOGGER.info("Reading the dataset from %s.", dataset)
        self.dataset = dataset if isinstance(dataset, str) else _to_container(dataset)
        super().__init__(name)
        self.hidden_attributes = BaseNodeBuilder.hidden_attributes | {"dataset"}

    def get_coordinates(self) -> torch.Tensor:
        """float32 (num_nodes, 2) coordinates of the nodes, in radians."""
        dataset = open_dataset(self.dataset)
        return self.reshape_coords(dataset.latitudes, dataset.longitudes)


def _to_container(config):
    try:  # pragma: no cover - omegaconf is not in this image
        from omegaconf import DictConfig, OmegaConf

        if isinstance(config, DictConfig):
            return OmegaConf.to_container(config)
    except ImportError:
        pass
    return dict(config) if hasattr(config, "items") else config


class TextNodes(BaseNodeBuilder):
    """Nodes from text file (from_file.py:66-93)."""

    def __init__(self, dataset, name: str, idx_lon: int = 0, idx_lat: int = 1) -> None:
        LOGGER.info("Reading the dataset from %s.", dataset)
        self.dataset = np.loadtxt(dataset)
        self.idx_lon = idx_lon
        self.idx_lat = idx_lat
        super().__init__(name)

    def get_coordinates(self) -> torch.Tensor:
        return self.reshape_coords(self.dataset[self.idx_lat, :], self.dataset[self.idx_lon, :])


class NPZFileNodes(BaseNodeBuilder):
    """Nodes from NPZ defined grids: ``<grid_definition_path>/grid-<resolution>.npz`` (from_file.py:96-151)."""

    def __init__(self, resolution: str, grid_definition_path: str, name: str) -> None:
        self.resolution = resolution
        self.grid_definition_path = grid_definition_path
        self.grid_definition = np.load(Path(self.grid_definition_path) / f"grid-{self.resolution}.npz")
        super().__init__(name)

    def get_coordinates(self) -> torch.Tensor:
        coords = self.reshape_coords(self.grid_definition["latitudes"], self.grid_definition["longitudes"])
        return coords


class LimitedAreaNPZFileNodes(NPZFileNodes):
    """Nodes from NPZ defined grids, limited to an area of interest (from_file.py:154-186)."""

    def __init__(
        self,
        resolution: str,
        grid_definition_path: str,
        reference_node_name: str,
        name: str,
        mask_attr_name: str | None = None,
        margin_radius_km: float = 100.0,
    ) -> None:
        self.area_mask_builder = KNNAreaMaskBuilder(reference_node_name, margin_radius_km, mask_attr_name)
        super().__init__(resolution, grid_definition_path, name)
        self.area_mask_builder = KNNAreaMaskBuilder(reference_node_name, margin_radius_km, mask_attr_name)

    def register_nodes(self, graph):
        self.area_mask_builder.fit(graph)
        return super().register_nodes(graph)

    def get_coordinates(self) -> torch.Tensor:
        coords = super().get_coordinates()
        LOGGER.info(
            "Limiting the processor mesh to a radius of %.2f km from the output mesh.",
            self.area_mask_builder.margin_radius_km,
        )
        area_mask = self.area_mask_builder.get_mask_device(coords).cpu()
        LOGGER.info("Dropping %d nodes from the processor mesh.", len(area_mask) - int(area_mask.sum()))
        return coords[area_mask]
