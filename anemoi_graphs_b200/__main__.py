"""``python -m anemoi_graphs_b200 create recipe.yaml graph.pt [--overwrite]`` - the ``anemoi-graphs create``
entry point (/root/reference/src/anemoi/graphs/commands/create.py:22-56) without the description step."""

from __future__ import annotations

import argparse
import logging
import sys

from .create import GraphCreator


def main(argv=None) -> int:
    parser = argparse.ArgumentParser(prog="anemoi-graphs")
    sub = parser.add_subparsers(dest="command", required=True)
    create = sub.add_parser("create", help="Create a graph from a recipe.")
    create.add_argument("--overwrite", action="store_true", help="Overwrite existing files. This will delete the target graph if it already exists.")
    create.add_argument("config", help="Configuration yaml file path defining the recipe to create the graph.")
    create.add_argument("save_path", help="Path to store the created graph. File format is torch .pt")
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(levelname)s %(name)s: %(message)s")
    GraphCreator(config=args.config).create(save_path=args.save_path, overwrite=args.overwrite)
    return 0


if __name__ == "__main__":
    sys.exit(main())
