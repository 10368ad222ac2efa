from topomax_b200.designs.design_parser import parse_design  # noqa: F401
