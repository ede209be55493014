"""Obstacle terrain: static world boxes appended to the MJCF (reference ``obstacles.py:16-57``).

Same draws as the reference for a given ``seed`` (Python ``random``: x, y, then yaw per box), same
geom attributes -- note ``size = (depth/2, length/2, height)``: ``height`` is used as a half-size and
the box is centred at z=0, so a "0.02 high" box protrudes 0.02 m above the floor.
"""

from __future__ import annotations

import math
import random
import xml.etree.ElementTree as ET
from typing import Tuple


def add_boxes_to_model(tree: ET.ElementTree, n_boxes: int, x_range: Tuple, y_range: Tuple, height: float = 0.02,
                       depth: float = 0.02, length: float = 3.0, group: str = "0", seed: int = 0) -> ET.ElementTree:
    worldbody = tree.getroot().find("worldbody")
    rnd = random.Random(seed)  # same Mersenne stream as random.seed(seed) + module-level draws
    half = f"{depth / 2.0} {length / 2.0} {height}"
    for i in range(n_boxes):
        x = rnd.uniform(x_range[0], x_range[1])
        y = rnd.uniform(y_range[0], y_range[1])
        yaw = rnd.uniform(-math.pi, math.pi)
        attrib = dict(name=f"box_geom_{i}", pos=f"{x} {y} 0", quat=f"{math.cos(yaw / 2)} 0 0 {math.sin(yaw / 2)}",
                      type="box", size=half, rgba="0.1 0.5 0.8 1", conaffinity="1", contype="1", condim="3", group=group)
        ET.SubElement(worldbody, "geom", attrib)
    return tree
