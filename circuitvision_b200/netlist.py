"""Netlist lines from the node table.

Behavioural mirror of `/root/reference/src/circuit_analyzer.py`:
  generate_netlist_from_nodes           :1607-1770
  stringify_line                        :1909-1927
  _get_terminal_nodes_relative_to_bbox  :1937-2030
with the class tables of :66-103 (netlist_map) and :128-130.  This is host-side bookkeeping on the small per-image
tables the device produced (node ids, attached component dicts, contour vertex arrays) — "netlist connectivity" in
BASELINE.json's sense.  Node centroids use the polygon moments OpenCV computes for an integer contour
(`cv2.moments`: a00 / a10 / a01 sums, m10/m00 and m01/m00 truncated by int()), evaluated here in exact integer
arithmetic so no OpenCV call is needed.  Parity: tests/test_netlist_cpu.py and tests/test_nodes_gpu.py compare the
text against the fixtures the unmodified reference produced (tests/golden/node_golden.npz, "netlist").
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

NETLIST_MAP = {  # circuit_analyzer.py:66-103
    "resistor": "R", "resistor.adjustable": "R",
    "capacitor": "C", "capacitor.unpolarized": "C", "capacitor.polarized": "C", "capacitor.adjustable": "C",
    "inductor": "L", "inductor.ferrite": "L",
    "diode": "D", "diode.light_emitting": "D", "diode.zener": "D",
    "transistor.bjt": "Q", "transistor.fet": "M",
    "voltage.ac": "V", "voltage.dc": "V", "voltage.battery": "V", "voltage.dependent": "E",
    "current.dc": "I", "current.ac": "I", "current.dependent": "G",
    "vss": "GND", "gnd": "0", "switch": "S",
    "integrated_circuit": "X", "integrated_circuit.voltage_regulator": "X", "operational_amplifier": "X",
    "thyristor": "Q", "transformer": "T", "varistor": "RV", "terminal": "N",
    "junction": "", "crossover": "", "explanatory": "", "text": "", "unknown": "UN",
}
VOLTAGE_CLASSES = frozenset(["voltage.dc", "voltage.ac", "transistor.bjt", "unknown"])  # :128
DIODE_CLASSES = frozenset(["diode", "diode.light_emitting", "diode.zener"])             # :129
CURRENT_SOURCE_CLASSES = frozenset(["current.dc", "current.dependent"])                 # :130
SKIPPED_CLASSES = ("text", "explanatory", "junction", "crossover")                       # :1651
_FLT_EPSILON = 1.1920928955078125e-07


def contour_centroid(contour):
    """(int(m10/m00), int(m01/m00)) of cv2.moments(contour) for an (N,1,2) / (N,2) integer polygon, or the first vertex
    when the polygon area vanishes (:1620-1628), or None for a missing contour."""
    if contour is None or len(contour) == 0:
        return None
    p = np.asarray(contour).reshape(-1, 2).astype(np.int64)
    x, y = p[:, 0], p[:, 1]
    xp, yp = np.roll(x, 1), np.roll(y, 1)  # previous vertex (the polygon closes on itself)
    cross = xp * y - x * yp
    a00 = int(cross.sum())
    if not abs(float(a00)) > _FLT_EPSILON:
        return (int(p[0, 0]), int(p[0, 1]))
    a10 = int((cross * (xp + x)).sum())
    a01 = int((cross * (yp + y)).sum())
    half, sixth = (0.5, 1.0 / 6.0) if a00 > 0 else (-0.5, -1.0 / 6.0)
    m00 = a00 * half
    return (int((a10 * sixth) / m00), int((a01 * sixth) / m00))


def order_terminal_nodes(component, direction, first_centroid, second_centroid, class_name, reason="UNKNOWN"):
    """:1937-2030 — which of the two node centroids is the component's primary terminal (positive pole / anode /
    arrow tail).  Without a semantic direction the SECOND one is primary (:1987)."""
    if not first_centroid or not second_centroid:
        return first_centroid, second_centroid
    cls = component.get("class", class_name)
    voltage_like, current_like = cls in VOLTAGE_CLASSES, cls in CURRENT_SOURCE_CLASSES
    arrow = current_like or (voltage_like and reason == "ARROW")
    signed = voltage_like and reason != "ARROW"
    if direction == "UNKNOWN" or not (arrow or signed or cls in DIODE_CLASSES):
        return second_centroid, first_centroid
    (ax, ay), (bx, by) = first_centroid, second_centroid
    tests = {"UP": ay < by, "DOWN": ay > by, "LEFT": ax < bx, "RIGHT": ax > bx}
    if direction not in tests:
        return first_centroid, second_centroid
    return (second_centroid, first_centroid) if tests[direction] else (first_centroid, second_centroid)


def generate_netlist_from_nodes(node_list, netlist_map=None):
    """:1607-1770 — one dict per placed component: component_type / component_num / node_1 / node_2 / value plus a
    deep copy of every field of the component's bbox dict."""
    nmap = NETLIST_MAP if netlist_map is None else netlist_map
    counters = {t: 1 for t in set(nmap.values()) if t}
    centroids = {n["id"]: contour_centroid(n.get("contour")) for n in node_list}
    seen, lines = set(), []
    for node in node_list:
        here = node["id"]
        for comp in node["components"]:
            cls, uid = comp.get("class"), comp.get("persistent_uid")
            if not uid or cls in SKIPPED_CLASSES or uid in seen:
                continue
            seen.add(uid)
            other = next((m["id"] for m in node_list
                          if m["id"] != here and any(c.get("persistent_uid") == uid for c in m["components"])), None)
            if cls == "terminal":  # still a terminal after the preliminary reclassification: one-node element to ground
                prefix, n1, n2 = nmap.get("terminal", "N"), here, "0"
            else:
                if other is None:
                    continue
                reason = comp.get("semantic_reason", "UNKNOWN")
                prefix = nmap.get(cls, "UN")
                if cls in VOLTAGE_CLASSES and reason == "ARROW":
                    prefix = "I"
                elif cls in CURRENT_SOURCE_CLASSES and reason == "SIGN":
                    prefix = "V"
                if not prefix:
                    continue
                c_here, c_other = centroids.get(here), centroids.get(other)
                if c_here is None or c_other is None:
                    first, second = here, other
                else:
                    primary, _ = order_terminal_nodes(comp, comp.get("semantic_direction", "UNKNOWN"), c_here, c_other,
                                                      cls, reason)
                    first, second = (here, other) if primary == c_here else (other, here)
                if cls in ("gnd", "vss"):
                    n1, n2 = (second if first == 0 else first), 0
                else:
                    n1, n2 = first, second
            if not prefix:
                continue
            num = counters.get(prefix, 1)
            counters[prefix] = num + 1
            line = {"component_type": prefix, "component_num": num, "node_1": n1, "node_2": n2, "value": "None"}
            line.update(deepcopy(comp))
            lines.append(line)
    return lines


def stringify_line(line) -> str:
    """:1909-1927 — SPICE text of one netlist line ('' for ground symbols and typeless entries)."""
    ctype = line.get("component_type")
    if line.get("class") == "gnd" or not ctype:
        return ""
    num, n1, n2 = line.get("component_num"), line.get("node_1"), line.get("node_2")
    if num is None or n1 is None or n2 is None:
        return ""
    return f"{ctype}{num} {n1} {n2} {line.get('value', 'None')}"


def netlist_text(node_list) -> str:
    """The fixture form: stringified lines joined by newlines (oracle/ref_loader.reference_node_analysis)."""
    return "\n".join(stringify_line(l) for l in generate_netlist_from_nodes(deepcopy(node_list)))
