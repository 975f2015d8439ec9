// Row f3 of the scope table: PDB text -> the padded per-structure arrays the hot path consumes.
//
// Host-only code (no kernels).  Replaces what the reference does through biotite + pandas in
// PDB.read_pdb / tidy_structure / PDB._initialize_lookup / PDB._compute_atom_xyz
// (protstruc/pdb.py:24-40, 55-151): first MODEL only, first alternate location per residue, non-standard
// residue names substituted, only the 20 canonical amino acids and standard heavy-atom names kept,
// residues in file order with UNK placeholders for numbering gaps inside a chain, chain index by first
// appearance, atom slot = position of the atom name in the residue type's 15-slot list.
// One pass over the text, no allocation beyond small fixed tables; thread-safe (no global state), so the
// Python side parses many files concurrently (ctypes releases the GIL).

#include <cstdlib>
#include <cstring>
#include <cmath>

#include "common.cuh"

namespace ps {

namespace {

constexpr int kSlots = 15;

struct ResidueType {
    const char* name;
    char one;
    const char* atoms[kSlots];
};

// Heavy-atom slot table (the AlphaFold-style atom14 + OXT layout the reference uses, general.py:149-171).
const ResidueType kTypes[20] = {
    {"ALA", 'A', {"N", "CA", "C", "O", "CB", "", "", "", "", "", "", "", "", "", "OXT"}},
    {"ARG", 'R', {"N", "CA", "C", "O", "CB", "CG", "CD", "NE", "CZ", "NH1", "NH2", "", "", "", "OXT"}},
    {"ASN", 'N', {"N", "CA", "C", "O", "CB", "CG", "OD1", "ND2", "", "", "", "", "", "", "OXT"}},
    {"ASP", 'D', {"N", "CA", "C", "O", "CB", "CG", "OD1", "OD2", "", "", "", "", "", "", "OXT"}},
    {"CYS", 'C', {"N", "CA", "C", "O", "CB", "SG", "", "", "", "", "", "", "", "", "OXT"}},
    {"GLN", 'Q', {"N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "NE2", "", "", "", "", "", "OXT"}},
    {"GLU", 'E', {"N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "OE2", "", "", "", "", "", "OXT"}},
    {"GLY", 'G', {"N", "CA", "C", "O", "", "", "", "", "", "", "", "", "", "", "OXT"}},
    {"HIS", 'H', {"N", "CA", "C", "O", "CB", "CG", "ND1", "CD2", "CE1", "NE2", "", "", "", "", "OXT"}},
    {"ILE", 'I', {"N", "CA", "C", "O", "CB", "CG1", "CG2", "CD1", "", "", "", "", "", "", "OXT"}},
    {"LEU", 'L', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "", "", "", "", "", "", "OXT"}},
    {"LYS", 'K', {"N", "CA", "C", "O", "CB", "CG", "CD", "CE", "NZ", "", "", "", "", "", "OXT"}},
    {"MET", 'M', {"N", "CA", "C", "O", "CB", "CG", "SD", "CE", "", "", "", "", "", "", "OXT"}},
    {"PHE", 'F', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "", "", "", "OXT"}},
    {"PRO", 'P', {"N", "CA", "C", "O", "CB", "CG", "CD", "", "", "", "", "", "", "", "OXT"}},
    {"SER", 'S', {"N", "CA", "C", "O", "CB", "OG", "", "", "", "", "", "", "", "", "OXT"}},
    {"THR", 'T', {"N", "CA", "C", "O", "CB", "OG1", "CG2", "", "", "", "", "", "", "", "OXT"}},
    {"TRP", 'W', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2", "OXT"}},
    {"TYR", 'Y', {"N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH", "", "", "OXT"}},
    {"VAL", 'V', {"N", "CA", "C", "O", "CB", "CG1", "CG2", "", "", "", "", "", "", "", "OXT"}},
};

// Non-standard -> standard residue names (the OpenMM/PDBFixer table the reference embeds,
// general.py:109-124), grouped by target.
struct Substitution {
    const char* target;
    const char* sources;  // space separated
};
const Substitution kSubstitutions[] = {
    {"ALA", "AIB ALM AYA BNN CHG CSD DAL DHA DNP FLA HAC MAA PRR TIH TPQ"},
    {"ARG", "ACL AGM ARM DAR HAR HMR"},
    {"ASN", "MEN"},
    {"ASP", "2AS ASA ASB ASK ASL ASQ BHD DAS DSP IAS"},
    {"CYS", "BCS BUC C5C C6C CAS CCS CEA CME CSO CSP CSS CSW CSX CY1 CY3 CYG CYM CYQ DCY EFC OCS PEC PR3 PYX SCH SCS SCY SHC SMC SOC"},
    {"GLN", "DGN"},
    {"GLU", "5HP CGU DGL GGL GMA PCA"},
    {"GLY", "GL3 GLZ GSC MPQ MSA NMC SAR"},
    {"HIS", "3AH DHI HIC HIP MHS NEM NEP"},
    {"ILE", "DIL IIL"},
    {"LEU", "BUG CLE DLE MLE NLE NLN NLP"},
    {"LYS", "ALY DLY KCX LLP LLY LYM LYZ SHR TRG"},
    {"MET", "CXM FME MSE OMT"},
    {"PHE", "DAH DPN HPQ PHI PHL"},
    {"PRO", "DPR HYP"},
    {"SER", "DSN MIS OAS SAC SEL SEP SET SVA"},
    {"THR", "ALO BMT DTH TPO"},
    {"TRP", "DTR HTR LTR TPL TRO"},
    {"TYR", "DTY IYR PAQ PTR STY TYB TYI TYQ TYS TYY"},
    {"VAL", "DIV DVA MVA"},
};

// Copies columns [begin, end) of a line without surrounding blanks into out (NUL terminated).
void field(const char* line, int line_len, int begin, int end, char* out, int out_cap) {
    if (end > line_len) end = line_len;
    while (begin < end && line[begin] == ' ') ++begin;
    while (end > begin && (line[end - 1] == ' ' || line[end - 1] == '\r')) --end;
    int n = end - begin;
    if (n < 0) n = 0;
    if (n > out_cap - 1) n = out_cap - 1;
    memcpy(out, line + begin, n);
    out[n] = '\0';
}

int type_index(const char* resname) {
    for (int t = 0; t < 20; ++t)
        if (strcmp(kTypes[t].name, resname) == 0) return t;
    return -1;
}

void substitute(char* resname) {
    if (type_index(resname) >= 0) return;
    for (const Substitution& s : kSubstitutions) {
        const char* p = s.sources;
        while (*p) {
            if (strncmp(p, resname, 3) == 0 && (p[3] == ' ' || p[3] == '\0') && strlen(resname) == 3) {
                strcpy(resname, s.target);
                return;
            }
            while (*p && *p != ' ') ++p;
            while (*p == ' ') ++p;
        }
    }
}

bool is_heavy_atom_name(const char* name) {
    if (!*name) return false;
    for (int t = 0; t < 20; ++t)
        for (int a = 0; a < kSlots; ++a)
            if (strcmp(kTypes[t].atoms[a], name) == 0) return true;
    return false;
}

int slot_of(int type, const char* atom) {
    for (int a = 0; a < kSlots; ++a)
        if (strcmp(kTypes[type].atoms[a], atom) == 0) return a;
    return -1;
}

struct Sink {
    int capacity;       // rows available in the output arrays (0 = counting only)
    float* xyz;
    uint8_t* mask;
    int32_t* chain_idx;
    char* chain_id;
    int32_t* resseq;
    char* icode;
    char* aa1;
    int rows = 0;
    char chains[64];
    int n_chains = 0;

    int chain_index(char c) {
        for (int k = 0; k < n_chains; ++k)
            if (chains[k] == c) return k;
        if (n_chains < 64) chains[n_chains] = c;
        return n_chains++;
    }
    // Opens a new row; returns its index (or -1 when only counting / out of capacity).
    int open_row(char chain, int number, char ins, char one) {
        const int idx = rows++;
        const int ci = chain_index(chain);
        if (idx >= capacity) return -1;
        const float nan = nanf("");
        for (int k = 0; k < kSlots * 3; ++k) xyz[idx * kSlots * 3 + k] = nan;
        memset(mask + idx * kSlots, 0, kSlots);
        chain_idx[idx] = ci;
        chain_id[idx] = chain;
        resseq[idx] = number;
        icode[idx] = ins;
        aa1[idx] = one;
        return idx;
    }
};

int parse(const char* text, long long len, Sink& sink) {
    bool seen_model = false, in_first_model = true;
    // identity of the residue currently being filled
    bool have_residue = false;
    char cur_chain = 0, cur_ins = 0, cur_name[4] = "";
    int cur_number = 0, cur_type = -1, cur_row = -1;
    char first_altloc = 0;
    // gap bookkeeping (reference pdb.py:92-121)
    bool have_chain = false;
    char gap_chain = 0;
    int gap_number = 0;

    long long pos = 0;
    while (pos < len) {
        const char* line = text + pos;
        const char* nl = static_cast<const char*>(memchr(line, '\n', len - pos));
        const int line_len = static_cast<int>(nl ? nl - line : len - pos);
        pos += line_len + 1;
        if (line_len < 6) continue;
        if (strncmp(line, "MODEL", 5) == 0) {
            if (seen_model) in_first_model = false;
            seen_model = true;
            continue;
        }
        if (strncmp(line, "ENDMDL", 6) == 0) {
            in_first_model = false;
            continue;
        }
        if (!in_first_model) continue;
        if (strncmp(line, "ATOM  ", 6) != 0 && strncmp(line, "HETATM", 6) != 0) continue;
        if (line_len < 54) continue;

        char atom[8], resname[8], num[8], coord[16];
        field(line, line_len, 12, 16, atom, sizeof atom);
        field(line, line_len, 17, 20, resname, sizeof resname);
        substitute(resname);
        const int type = type_index(resname);
        if (type < 0 || !is_heavy_atom_name(atom)) continue;
        const char altloc = line[16] == ' ' ? 0 : line[16];
        const char chain = line[21];
        const char ins = line[26] == ' ' ? 0 : line[26];
        field(line, line_len, 22, 26, num, sizeof num);
        const int number = atoi(num);

        const bool same = have_residue && chain == cur_chain && number == cur_number && ins == cur_ins &&
                          strcmp(resname, cur_name) == 0;
        if (!same) {
            // a new residue starts: fill numbering gaps inside the chain with UNK rows first
            if (!have_chain || gap_chain != chain) {
                gap_chain = chain;
                gap_number = number;
                have_chain = true;
            }
            while (gap_number + 1 < number) {
                sink.open_row(gap_chain, gap_number + 1, ins, 'X');
                ++gap_number;
            }
            cur_row = sink.open_row(chain, number, ins, kTypes[type].one);
            gap_chain = chain;
            gap_number = number;
            have_residue = true;
            cur_chain = chain;
            cur_number = number;
            cur_ins = ins;
            strcpy(cur_name, resname);
            cur_type = type;
            first_altloc = 0;
        }
        if (altloc) {  // keep only the first alternate location seen in this residue
            if (!first_altloc) first_altloc = altloc;
            if (altloc != first_altloc) continue;
        }
        const int slot = slot_of(cur_type, atom);
        if (slot < 0 || cur_row < 0) continue;
        float* dst = sink.xyz + (static_cast<long long>(cur_row) * kSlots + slot) * 3;
        for (int k = 0; k < 3; ++k) {
            field(line, line_len, 30 + 8 * k, 38 + 8 * k, coord, sizeof coord);
            dst[k] = static_cast<float>(strtod(coord, nullptr));
        }
        sink.mask[cur_row * kSlots + slot] = 1;
    }
    return sink.rows;
}

}  // namespace

int host_pdb_parse_impl(const char* text, long long len, int capacity, float* xyz, uint8_t* mask,
                        int32_t* chain_idx, char* chain_id, int32_t* resseq, char* icode, char* aa1,
                        int* n_residues) {
    PS_REQUIRE(text != nullptr && len >= 0, PS_ERR_NULL_POINTER, "pdb_parse: text is NULL");
    PS_REQUIRE(n_residues != nullptr, PS_ERR_NULL_POINTER, "pdb_parse: n_residues is NULL");
    PS_REQUIRE(capacity == 0 || (xyz && mask && chain_idx && chain_id && resseq && icode && aa1),
               PS_ERR_NULL_POINTER, "pdb_parse: output arrays are required when capacity > 0");
    Sink sink{};
    sink.capacity = capacity;
    sink.xyz = xyz;
    sink.mask = mask;
    sink.chain_idx = chain_idx;
    sink.chain_id = chain_id;
    sink.resseq = resseq;
    sink.icode = icode;
    sink.aa1 = aa1;
    *n_residues = parse(text, len, sink);
    PS_REQUIRE(capacity == 0 || *n_residues <= capacity, PS_ERR_BAD_SHAPE,
               "pdb_parse: %d residues do not fit in the %d rows provided", *n_residues, capacity);
    return PS_OK;
}

}  // namespace ps
