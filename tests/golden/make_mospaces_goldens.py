"""Extract the hand-written MoSpaces expectations of the reference's C++ test
(libadcc_src/tests/MoSpacesTest.cc:31-552, mock SCF data libadcc_src/tests/HFSolutionMock.hh:38-44)
into tests/golden/mospaces_goldens.json.  Runs in the build container only (reads /root/reference).
"""
import json
import os
import re

SRC = "/root/reference/libadcc_src/tests/MoSpacesTest.cc"
MOCK = "/root/reference/libadcc_src/tests/HFSolutionMock.hh"
HERE = os.path.dirname(os.path.abspath(__file__))


def numbers(text):
    return [float(x) if ("." in x or "e" in x.lower()) else int(x)
            for x in re.findall(r"-?\d+\.?\d*(?:[eE][+-]?\d+)?", text)]


def main():
    mock = open(MOCK).read()
    occ = numbers(re.search(r"exposed_occupation\s*=\s*\{([^}]*)\}", mock).group(1))
    orben = numbers(re.search(r"exposed_orben\s*=\s*\{([^}]*)\}", mock, re.S).group(1))
    text = open(SRC).read()
    pieces = re.split(r'SECTION\("([^"]*)"\)', text)[1:]
    cases = []
    for name, body in zip(pieces[0::2], pieces[1::2]):
        bs = int(re.search(r"block_size\s*=\s*(\d+)", body).group(1))
        args = re.search(r"MoSpaces mo\(tested, adcmem_ptr,\s*\{([^}]*)\},\s*\{([^}]*)\},\s*"
                         r"\{([^}]*)\}\)", body)
        case = {"name": name, "block_size": bs,
                "core_orbitals": numbers(args.group(1)), "frozen_core": numbers(args.group(2)),
                "frozen_virtual": numbers(args.group(3)),
                "map_index_hf_provider": {}, "map_block_start": {}, "map_block_spin": {},
                "n_orbs_alpha": {}}
        for key in ("map_index_hf_provider", "map_block_start"):
            for m in re.finditer(key + r'\["(\w+)"\]\s*==\s*std::vector<size_t>\{([^}]*)\}', body):
                case[key][m.group(1)] = numbers(m.group(2))
        for m in re.finditer(r'map_block_spin\["(\w+)"\]\s*==\s*std::vector<char>\{([^}]*)\}', body):
            case["map_block_spin"][m.group(1)] = re.findall(r"'(\w)'", m.group(2))
        for m in re.finditer(r'n_orbs_alpha\("(\w+)"\)\s*==\s*(\d+)', body):
            case["n_orbs_alpha"][m.group(1)] = int(m.group(2))
        for key in ("subspaces_occupied", "subspaces_virtual", "subspaces"):
            m = re.search(r"mo\." + key + r"\s*==\s*std::vector<std::string>\{([^}]*)\}", body)
            case[key] = re.findall(r'"(\w+)"', m.group(1))
        cases.append(case)
    out = {"occupation_f": occ, "orben_f": orben, "cases": cases}
    with open(os.path.join(HERE, "mospaces_goldens.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()


def motrans_cases():
    """Scalar index-translation expectations of libadcc_src/tests/MoIndexTranslationTest.cc."""
    text = open("/root/reference/libadcc_src/tests/MoIndexTranslationTest.cc").read()
    token = re.compile(
        r"block_size\s*=\s*(?P<bs>\d+)"
        r"|new MoSpaces\(tested, adcmem_ptr,\s*\{(?P<c>[^}]*)\},\s*\{(?P<fc>[^}]*)\},\s*\{(?P<fv>[^}]*)\}\)"
        r'|MoIndexTranslation motrans\(mospaces_ptr,\s*"(?P<space>\w+)"\)'
        r"|(?<!_THROWS)\(motrans\.(?P<fn>block_index_of|block_index_spatial_of|inblock_index_of|"
        r"spin_of|hf_provider_index_of)\(\{(?P<idx>[^}]*)\}\)\s*==\s*"
        r'(?:vecsize\{(?P<vec>[^}]*)\}|"(?P<spin>\w+)")\)')
    # range mappings: ``std::vector<RangeMapping> ref_N{ {from pairs},{to pairs} ... };`` followed by
    # ``CHECK(motrans.map_range_to_hf_provider({ranges}) == ref_N);``
    ref_def = re.compile(r"std::vector<RangeMapping>\s*(ref_\d+)\s*\{(.*?)\};", re.S)
    ref_use = re.compile(r"map_range_to_hf_provider\(\{(.*?)\}\)\s*==\s*(ref_\d+)\)", re.S)
    events = sorted([(m.start(), "tok", m) for m in token.finditer(text)]
                    + [(m.start(), "def", m) for m in ref_def.finditer(text)]
                    + [(m.start(), "use", m) for m in ref_use.finditer(text)], key=lambda e: e[0])
    out, bs, args, cur, refs = [], None, None, None, {}
    for _, kind, m in events:
        if kind == "def":
            refs[m.group(1)] = numbers(m.group(2))
            continue
        if kind == "use":
            ndim = len(cur["space"]) // 2
            rng = numbers(m.group(1))
            flat = refs[m.group(2)]
            assert len(rng) == 2 * ndim and len(flat) % (4 * ndim) == 0
            pairs = lambda v: [v[2 * k:2 * k + 2] for k in range(len(v) // 2)]      # noqa: E731
            maps = [[pairs(flat[o:o + 2 * ndim]), pairs(flat[o + 2 * ndim:o + 4 * ndim])]
                    for o in range(0, len(flat), 4 * ndim)]
            cur.setdefault("range_checks", []).append([pairs(rng), maps])
            continue
        if m.group("bs"):
            bs = int(m.group("bs"))
        elif m.group("c") is not None:
            args = [numbers(m.group(k)) for k in ("c", "fc", "fv")]
        elif m.group("space"):
            cur = {"block_size": bs, "core_orbitals": args[0], "frozen_core": args[1],
                   "frozen_virtual": args[2], "space": m.group("space"), "checks": []}
            out.append(cur)
        elif m.group("fn"):
            expect = m.group("spin") if m.group("spin") else numbers(m.group("vec"))
            cur["checks"].append([m.group("fn"), numbers(m.group("idx")), expect])
    return out


if __name__ == "__main__":
    cases = motrans_cases()
    path = os.path.join(HERE, "mospaces_goldens.json")
    data = json.load(open(path))
    data["motrans"] = cases
    with open(path, "w") as fp:
        json.dump(data, fp, indent=1)
    print(len(cases), "index-translation spaces,", sum(len(c["checks"]) for c in cases), "checks")
