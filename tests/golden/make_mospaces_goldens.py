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
    out, bs, args, cur = [], None, None, None
    for m in token.finditer(text):
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
