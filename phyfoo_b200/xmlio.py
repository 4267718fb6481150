"""Jstacs-XML model files of the reference, read and written without a JVM.

The reference stores trained models as the text Jstacs' `XMLParser` produces (`jstacs!/de/jstacs/io/XMLParser.java:127-205`):
`<tag><className>java.lang.Integer</className>12</tag>` for a value, and for an array
`<tag><className>[[D</className><length>2</length><pos val="0">...</pos><pos val="1">...</pos></tag>`
where the elements of an array of primitives / strings carry no class name.  What is stored for the models of the path:

  PhyloBayesModel  (models/PhyloBayesModel.java:291-309):  <_bnh> ... </_bnh> <_order> <modelLength> <alphabets>
  PhyloBackground  (models/PhyloBackground.java:330-346):  <_bnh> ... </_bnh> <order> <alphabets>
  BayesNetHandler  (bayesNet/BayesNetHandler.java:288-357): <numberOfTrees> <connectionTable> and per position
                   <VirtualTree i> <evolModel> <statDistr> <dimension> </evolModel> <TEMPERATURE> <newickString>
  FS81alpha        (evolution/FS81alpha.java:109-128), Newick with branch lengths rounded to three decimals
                   (bayesNet/VirtualTree.java:208-234)
  mixtures         (jstacs AbstractMixtureModel.java:1628-1678): SingleHiddenMotifMixture / StrandModel wrap <models>, <weights>

Reading follows the reference's reload semantics: a reloaded model is always F81-alpha (BayesNetHandler.java:325,337, SURVEY.md
appendix B-4) and its branch lengths are the three-decimal values of the file.  No model file ships with the reference and there
is no JVM here, so the format is restated from the serialisation code and checked by round trips and hand-assembled files
(tests/test_xmlio.py); files written here have not been loaded by the Java side.  Mixture files (SingleHiddenMotifMixture,
StrandModel) are written as well (shm_to_xml), with the termination condition in the loader's legacy <epsilon> form.
"""
import re
from decimal import Decimal

import numpy as np

from .newick import parse_newick

_ESC = [("&", "&amp;"), ('"', "&quot;"), ("'", "&apos;"), (">", "&gt;"), ("<", "&lt;")]     # XMLParser.java:811-817


def _escape(s):
    for a, b in _ESC:
        s = s.replace(a, b)
    return s


def _unescape(s):
    for a, b in _ESC:
        s = s.replace(b, a)
    return s


# ---------------------------------------------------------------------------------------------------------------------------
# parsing: the text is tag soup with mixed content, not strict XML -- build a small tree
# ---------------------------------------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("tag", "attrs", "children", "text")

    def __init__(self, tag, attrs):
        self.tag, self.attrs, self.children, self.text = tag, attrs, [], ""

    def child(self, tag, **attrs):
        """First direct child with this tag (and attribute values), or None."""
        for c in self.children:
            if c.tag == tag and all(c.attrs.get(k) == str(v) for k, v in attrs.items()):
                return c
        return None

    def find(self, tag):
        """First node with this tag in document order (what XMLParser.extractForTag finds: XMLParser.java:380-420)."""
        for c in self.children:
            if c.tag == tag:
                return c
            r = c.find(tag)
            if r is not None:
                return r
        return None


_TAG = re.compile(r"<(/?)([^\s<>/]+)([^<>]*)>")
_ATTR = re.compile(r"([^\s=]+)\s*=\s*\"([^\"]*)\"")


def parse(text):
    """Parse Jstacs XML text into a tree; returns the (tagless) root node."""
    root = Node("", {})
    stack = [root]
    pos = 0
    for m in _TAG.finditer(text):
        stack[-1].text += text[pos:m.start()]
        pos = m.end()
        closing, tag, rest = m.group(1), m.group(2), m.group(3)
        if closing:
            if len(stack) == 1 or stack[-1].tag != tag:
                raise ValueError(f"Malformed XML: Found unexpected end tag: </{tag}>.")
            stack.pop()
        else:
            n = Node(tag, dict(_ATTR.findall(rest)))
            stack[-1].children.append(n)
            stack.append(n)
    if len(stack) != 1:
        raise ValueError(f"Malformed XML: <{stack[-1].tag}> is not closed")
    return root


_INT = {"java.lang.Integer", "java.lang.Byte", "java.lang.Short", "java.lang.Long", "int", "byte", "short", "long"}
_FLT = {"java.lang.Double", "java.lang.Float", "double", "float"}
_PRIM = {"D": "double", "F": "float", "I": "int", "J": "long", "S": "short", "B": "byte", "Z": "boolean", "C": "char"}


def _simple(cls, text):
    text = text.strip()
    if cls in _INT:
        return int(text)
    if cls in _FLT:
        return float(text)
    if cls in ("java.lang.Boolean", "boolean"):
        return text == "true"
    if cls in ("java.lang.String", "java.lang.Character", "char"):
        return _unescape(text)
    return None


def value(node, cls=None):
    """Decode what XMLParser.appendObjectWithTags wrote into `node`.  Values of simple classes become Python scalars, arrays
    become lists (numpy arrays for numeric ones), anything else (a Storable) is returned as its node."""
    if node is None:
        raise KeyError("tag not found")
    cn = node.child("className")
    if cn is not None:
        cls = cn.text.strip()
    if cls is None:
        raise ValueError(f"<{node.tag}> carries no class information")
    if not node.children and node.text.strip() == "null":
        return None
    if cls.startswith("["):
        comp = cls[1:]
        if comp.startswith("L") and comp.endswith(";"):
            comp = comp[1:-1]
        elif not comp.startswith("["):
            comp = _PRIM[comp]
        n = int(node.child("length").text.strip())
        items = []
        for i in range(n):
            el = node.child("pos", val=i)
            if el is None:
                raise ValueError(f"<{node.tag}>: array element {i} is missing")
            items.append(value(el, comp))
        if items and all(isinstance(x, (int, float, np.ndarray)) and not isinstance(x, bool) for x in items):
            try:
                return np.array(items)
            except ValueError:
                return items
        return items
    v = _simple(cls, node.text)
    return node if v is None else v


# ---------------------------------------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------------------------------------
def jdouble(x):
    """java.lang.Double.toString: decimal notation for 1e-3 <= |x| < 1e7, otherwise d.dddE[-]n; shortest round-trip digits."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    if 1e-3 <= abs(x) < 1e7:
        r = repr(x)
        if "e" in r or "E" in r:
            r = format(Decimal(r), "f")
        return r if "." in r else r + ".0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    ds = "".join(map(str, digits))
    e = exp + len(ds) - 1                                   # decimal exponent of the leading digit
    ds = ds.rstrip("0") or "0"
    return ("-" if sign else "") + ds[0] + "." + (ds[1:] or "0") + "E" + str(e)


class JValue:
    """A value with its Java class: JValue('java.lang.Byte', 1), JValue('[[D', array), JValue('models.PhyloBayesModel', xml_text)."""

    def __init__(self, cls, v):
        self.cls, self.v = cls, v


def _scalar_text(cls, v):
    if cls in _INT:
        return str(int(v))
    if cls in _FLT:
        return jdouble(v)
    if cls in ("java.lang.Boolean", "boolean"):
        return "true" if v else "false"
    if cls in ("java.lang.String",):
        return _escape(str(v))
    return None


def append_object(jv, tag, attributes=None, class_info=True):
    """XMLParser.appendObjectWithTagsAndAttributes (XMLParser.java:170-205) for a JValue."""
    out = ["<" + tag + ("" if attributes is None else " " + attributes) + ">"]
    if jv is None or jv.v is None:
        out.append("null")
    else:
        cls = jv.cls
        if class_info:
            out.append("<className>" + _escape(cls) + "</className>")
        if cls.startswith("["):
            comp = cls[1:]
            simple_comp = False
            if comp.startswith("L") and comp.endswith(";"):
                comp = comp[1:-1]
                simple_comp = comp == "java.lang.String"
            elif not comp.startswith("["):
                comp, simple_comp = _PRIM[comp], True
            items = list(jv.v)
            out.append("<length>" + str(len(items)) + "</length>")
            for i, it in enumerate(items):
                el = it if isinstance(it, JValue) else JValue(comp, it)
                out.append(append_object(el, "pos", f'val="{i}"', class_info and not simple_comp))
        else:
            t = _scalar_text(cls, jv.v)
            out.append(str(jv.v) if t is None else t)          # a Storable's own toXML() text
    out.append("</" + tag + ">")
    return "".join(out)


def add_tags(xml, tag):
    return "<" + tag + ">\n" + xml + "</" + tag + ">\n"               # XMLParser.java:108-111


def _alphabets_xml():
    """AlphabetContainer(new UnobservableDNAAlphabet()) (training/TrainModel.java:53): AlphabetContainer.java:973-981,
    DiscreteAlphabet.java:210-216."""
    alpha = add_tags(append_object(JValue("[Ljava.lang.String;", ["A", "C", "G", "T", "-"]), "symbols") +
                     append_object(JValue("java.lang.Boolean", True), "caseInsensitive"), "DiscreteAlphabet")
    cont = add_tags(append_object(JValue("[Lde.jstacs.data.Alphabet;",
                                         [JValue("de.jstacs.data.alphabets.UnobservableDNAAlphabet", alpha)]), "Alphabets"),
                    "AlphabetContainer")
    return append_object(JValue("de.jstacs.data.AlphabetContainer", cont), "alphabets")


def newick_string(tree, branch):
    """bayesNet/VirtualTree.java:208-234: children in order, `child:round(length, 3)`, leaves by name."""
    kids = [[] for _ in range(tree.n_nodes)]
    for k in range(1, tree.n_nodes):
        kids[tree.parent[k]].append(k)

    def rec(n):
        if not kids[n]:
            return tree.labels[n]
        return "(" + ",".join(rec(c) + ":" + jdouble(round(float(branch[c]) * 1000.0) / 1000.0) for c in kids[n]) + ")"
    return rec(0) + ";"


def _bnh_xml(model):
    n = model.n_trees
    ct = np.zeros((n, n), np.int32)
    if model.order >= 1:
        for t in range(1, n):
            ct[t - 1, t] = 1                                          # tree t-1 is the "horizontal" parent of tree t
    xml = append_object(JValue("java.lang.Integer", n), "numberOfTrees")
    xml += append_object(JValue("[[I", [JValue("[I", row) for row in ct]), "connectionTable")
    for t in range(n):
        pi = np.asarray(model._pi[t], np.float64).reshape(-1, 4)
        s = add_tags(append_object(JValue("[[D", [JValue("[D", row) for row in pi]), "statDistr") +
                     append_object(JValue("java.lang.Integer", pi.shape[0]), "dimension"), "evolModel")
        s += append_object(JValue("java.lang.Double", 1.0), "TEMPERATURE")
        s += append_object(JValue("java.lang.String", newick_string(model.tree, model._branch[t])), "newickString")
        xml += add_tags(s, "VirtualTree" + str(t))
    return xml


def phylo_bayes_model_to_xml(model):
    """models/PhyloBayesModel.java:291-299."""
    xml = add_tags(_bnh_xml(model), "_bnh")
    xml += append_object(JValue("java.lang.Byte", model.order), "_order")
    xml += append_object(JValue("java.lang.Integer", model.length), "modelLength")
    return xml + _alphabets_xml()


def phylo_background_to_xml(model):
    """models/PhyloBackground.java:330-336."""
    xml = add_tags(_bnh_xml(model), "_bnh")
    xml += append_object(JValue("java.lang.Byte", model.order), "order")
    return xml + _alphabets_xml()


# ---------------------------------------------------------------------------------------------------------------------------
# mixture writers (jstacs AbstractMixtureModel.toXML, :1628-1678)
_ENUM_ALGORITHM = "de.jstacs.models.mixture.AbstractMixtureModel$Algorithm"
_ENUM_PARAMETERIZATION = "de.jstacs.models.mixture.AbstractMixtureModel$Parameterization"


def _enum_xml(cls, name, tag):
    """XMLParser.java:190-191: an enum is its class name and <name>CONSTANT</name>."""
    return "<" + tag + "><className>" + _escape(cls) + "</className><name>" + name + "</name></" + tag + ">"


def _mixture_to_xml(simple_name, length, models, optimize, estimate, hyper, weights, trained, best, epsilon, alpha=1.0, further=""):
    """The field order of AbstractMixtureModel.toXML.  The termination condition is written in the form the loader still accepts
    from older files -- <epsilon> for a SmallDifferenceOfFunctionEvaluationsCondition (AbstractMixtureModel.java:1713-1714) --
    instead of the nested ParameterSet of a CombinedCondition: PhyFoo's pipelines replace the loaded model's condition before
    they train on (training/TrainingUtil.java:56-68, PostTrainModel.java:90), and scoring does not use it."""
    x = append_object(JValue("java.lang.Integer", length), "length")
    x += append_object(JValue("java.lang.Integer", 2), "dimension")
    x += append_object(JValue("java.lang.Integer", 1), "starts")                          # models/ModelUtil.java:101,114
    x += append_object(JValue("java.lang.Boolean", estimate), "estimateComponentProbs")
    x += append_object(JValue("[D", list(np.asarray(hyper, np.float64))), "componentHyperParams")
    x += append_object(JValue("[Lde.jstacs.models.Model;", models), "models")
    x += append_object(JValue("[Z", optimize), "optimizeModel")
    x += append_object(JValue("java.lang.Boolean", trained), "algorithmHasBeenRun")
    x += append_object(JValue("[D", list(np.asarray(weights, np.float64))), "weights")
    x += _enum_xml(_ENUM_ALGORITHM, "EM", "algorithm")
    x += append_object(JValue("java.lang.Double", alpha), "alpha")
    x += append_object(JValue("java.lang.Double", epsilon), "epsilon")
    x += _enum_xml(_ENUM_PARAMETERIZATION, "LAMBDA", "parametrization")
    x += append_object(JValue("java.lang.Double", float("-inf") if best is None else best), "best")
    return add_tags(x + further, simple_name)


def _component_jvalue(model):
    from .models import PhyloBackground, PhyloBayesModel, StrandModel
    if isinstance(model, StrandModel):
        return JValue("de.jstacs.models.mixture.StrandModel", strand_model_to_xml(model))
    if isinstance(model, PhyloBayesModel):
        return JValue("models.PhyloBayesModel", phylo_bayes_model_to_xml(model))
    if isinstance(model, PhyloBackground):
        return JValue("models.PhyloBackground", phylo_background_to_xml(model))
    raise NotImplementedError(f"{type(model).__name__} in a mixture file")


def strand_model_to_xml(sm, trained=True, best=None):
    """jstacs StrandModel as models/ModelUtil.java:98-108 builds it: one inner model, two components (forward, backward)."""
    return _mixture_to_xml("StrandModel", sm.getLength(), [_component_jvalue(sm.model)], [True], sm.estimateComponentProbs, sm.hyper,
                           sm.weights, trained, best, 1e-4, sm.alpha)


def shm_to_xml(shm, trained=True, best=None, epsilon=1e-4):
    """SingleHiddenMotifMixture.toXML: AbstractMixtureModel's fields, then the position prior of HiddenMotifMixture.java:208-213
    (a UniformPositionPrior, PositionPrior.java:121-127), framed by the class's simple name.  models = {motif (possibly inside a
    StrandModel), flanking}; only the motif is optimised (trainOnlyMotifModel = true, models/ModelUtil.java:112)."""
    motif = shm.strand if shm.strand is not None else shm.motif
    prior = add_tags(append_object(JValue("java.lang.Integer", shm.motif.getLength()), "motifLength"), "UniformPositionPrior")
    further = append_object(JValue("de.jstacs.models.mixture.motif.positionprior.UniformPositionPrior", prior), "posPrior")
    return _mixture_to_xml("SingleHiddenMotifMixture", 0, [_component_jvalue(motif), _component_jvalue(shm.flanking)], [True, False],
                           shm.estimateComponentProbs, shm.hyper, shm.weights, trained, best, epsilon, 1.0, further)


def save_shm(path, shm, **kw):
    with open(path, "w") as f:
        f.write(shm_to_xml(shm, **kw))


# ---------------------------------------------------------------------------------------------------------------------------
# model readers
# ---------------------------------------------------------------------------------------------------------------------------
def _read_bnh(node):
    """-> (n_trees, order, [newick per tree], [statDistr per tree], [temperature per tree])   (BayesNetHandler.java:313-357)"""
    n = value(node.child("numberOfTrees"))
    ct = np.asarray(value(node.child("connectionTable")))
    newicks, pis, temps = [], [], []
    for t in range(n):
        vt = node.child("VirtualTree" + str(t))
        if vt is None:
            raise ValueError(f"VirtualTree{t} is missing")
        newicks.append(value(vt.child("newickString")))
        temps.append(value(vt.child("TEMPERATURE")) if vt.child("TEMPERATURE") is not None else 1.0)
        ev = vt.child("evolModel")
        pi = np.asarray(value(ev.child("statDistr")), np.float64).reshape(-1, 4)
        dim = value(ev.child("dimension"))
        if pi.shape[0] != dim:
            raise ValueError(f"VirtualTree{t}: statDistr has {pi.shape[0]} rows, dimension says {dim}")
        pis.append(pi)
    return n, (1 if ct.size and ct.any() else 0), newicks, pis, temps


def _fill(model, newicks, pis, temps):
    if any(abs(t - 1.0) > 1e-12 for t in temps):
        raise NotImplementedError("TEMPERATURE != 1 (optimizing/TemperatureOptimizer.java) is not supported")
    model.reinitEdgelengths(*newicks)
    for t, pi in enumerate(pis):
        if pi.shape != np.asarray(model._pi[t]).reshape(-1, 4).shape:
            raise ValueError(f"position {t}: {pi.shape[0]} context rows in the file, the model has "
                             f"{np.asarray(model._pi[t]).reshape(-1, 4).shape[0]}")
        model.setCondProb(pi, t)
    return model


def read_phylo_bayes_model(src, ctx=None):
    """PhyloBayesModel(StringBuffer) (models/PhyloBayesModel.java:300-309): text or an already parsed node."""
    from .models import PhyloBayesModel
    node = parse(src) if isinstance(src, str) else src
    n, order_ct, newicks, pis, temps = _read_bnh(node.child("_bnh"))
    order = value(node.child("_order"))
    length = value(node.child("modelLength"))
    if length != n:
        raise ValueError(f"modelLength {length} but {n} trees")
    return _fill(PhyloBayesModel(length, order, newicks[0], ctx), newicks, pis, temps)


def read_phylo_background(src, ctx=None):
    """PhyloBackground(StringBuffer) (models/PhyloBackground.java:338-346)."""
    from .models import PhyloBackground
    node = parse(src) if isinstance(src, str) else src
    n, order_ct, newicks, pis, temps = _read_bnh(node.child("_bnh"))
    order = value(node.child("order"))
    if n != order + 1:
        raise ValueError(f"order {order} but {n} trees")
    return _fill(PhyloBackground(order, newicks[0], ctx), newicks, pis, temps)


def _read_component(el, ctx):
    """One entry of a mixture's <models> array -> (model, kind)."""
    from .models import StrandModel
    cls = el.child("className").text.strip()
    if cls.endswith("StrandModel"):
        sm = el.child("StrandModel")
        inner = sm.child("models").child("pos", val=0)
        motif, _ = _read_component(inner, ctx)
        w = np.asarray(value(sm.child("weights")), np.float64)
        est = value(sm.child("estimateComponentProbs"))
        hyper = value(sm.child("componentHyperParams"))
        return StrandModel(motif, float(w[0]), (1.0, 1.0) if hyper is None else tuple(np.asarray(hyper, float)), est), "strand"
    if cls.endswith("PhyloBayesModel"):
        return read_phylo_bayes_model(el, ctx), "motif"
    if cls.endswith("PhyloBackground"):
        return read_phylo_background(el, ctx), "background"
    raise NotImplementedError(f"model class {cls} in a mixture file")


def read_shm(text, ctx=None):
    """SingleHiddenMotifMixture(StringBuffer) as tools/Classification and models/ModelUtil.java:228-237 load it: the motif
    (possibly inside a StrandModel), the flanking model and the component weights."""
    from .models import SingleHiddenMotifMixture
    root = parse(text)
    shm = root.child("SingleHiddenMotifMixture") or root.find("SingleHiddenMotifMixture")
    if shm is None:
        raise ValueError("no <SingleHiddenMotifMixture> in the file")
    models = shm.child("models")
    motif, _ = _read_component(models.child("pos", val=0), ctx)
    flank, kind = _read_component(models.child("pos", val=1), ctx)
    if kind != "background":
        raise ValueError("the second model of the mixture is not a background model")
    w = np.asarray(value(shm.child("weights")), np.float64)
    est = value(shm.child("estimateComponentProbs"))
    hyper = value(shm.child("componentHyperParams"))
    out = SingleHiddenMotifMixture(motif, flank, zoops=None, componentHyperParams=(1.0, 1.0) if hyper is None else tuple(np.asarray(hyper, float)))
    out.weights = w.copy()
    out.estimateComponentProbs = bool(est)
    return out


def load_background(path, ctx=None):
    """models/ModelUtil.java:204-220."""
    text = open(path).read()
    if "<_bnh" not in text:
        raise NotImplementedError("AlignmentBasedBGModel files are not supported")
    return read_phylo_background(text, ctx)


def load_shm(path, ctx=None):
    return read_shm(open(path).read(), ctx)
