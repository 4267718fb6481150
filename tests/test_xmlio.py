"""Jstacs-XML model files (phyfoo_b200/xmlio.py).  No model file ships with the reference and no JVM
exists here, so the checks are: the low-level format against text typed in from XMLParser.appendObjectWithTags' rules
(jstacs XMLParser.java:170-205), Double.toString formatting, round trips of the models, the reference's three-decimal branch
rounding on save, and a mixture file assembled the way AbstractMixtureModel.toXML (jstacs :1628-1678) nests its parts."""
import numpy as np
import pytest

from phyfoo_b200 import xmlio
from phyfoo_b200.xmlio import JValue, add_tags, append_object


def test_low_level_format():
    assert append_object(JValue("java.lang.Integer", 12), "numberOfTrees") == \
        "<numberOfTrees><className>java.lang.Integer</className>12</numberOfTrees>"
    a = append_object(JValue("[[D", [JValue("[D", [0.25, 0.5]), JValue("[D", [1e-5, 3.0])]), "statDistr")
    assert a == ('<statDistr><className>[[D</className><length>2</length>'
                 '<pos val="0"><className>[D</className><length>2</length><pos val="0">0.25</pos><pos val="1">0.5</pos></pos>'
                 '<pos val="1"><className>[D</className><length>2</length><pos val="0">1.0E-5</pos><pos val="1">3.0</pos></pos>'
                 '</statDistr>')
    back = xmlio.value(xmlio.parse(a).child("statDistr"))
    assert back.shape == (2, 2) and back[1, 0] == 1e-5
    s = append_object(JValue("java.lang.String", "(A:0.1,B:0.2)<&>;"), "newickString")
    assert "&lt;&amp;&gt;" in s and xmlio.value(xmlio.parse(s).child("newickString")) == "(A:0.1,B:0.2)<&>;"
    assert append_object(None, "best") == "<best>null</best>"
    assert xmlio.value(xmlio.parse("<x><className>java.lang.Boolean</className>true</x>").child("x")) is True
    assert add_tags("abc", "_bnh") == "<_bnh>\nabc</_bnh>\n"
    with pytest.raises(ValueError):
        xmlio.parse("<a><b></a></b>")


def test_java_double_to_string():
    for x, s in [(0.25, "0.25"), (1.0, "1.0"), (0.001, "0.001"), (1e-4, "1.0E-4"), (1.5e-7, "1.5E-7"), (1e7, "1.0E7"),
                 (12345678.9, "1.23456789E7"), (9999999.0, "9999999.0"), (-0.2, "-0.2"), (0.0, "0.0"), (100.0, "100.0"),
                 (1.0 / 3.0, "0.3333333333333333"), (2.5e-10, "2.5E-10")]:
        assert xmlio.jdouble(x) == s, (x, xmlio.jdouble(x), s)
        assert float(xmlio.jdouble(x)) == x


def _models():
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel
    newick = "((A:0.1234,B:0.3):0.15,(C:0.2,(D:0.05049,E:0.4):0.25):0.1,F:0.3)"
    fg0 = PhyloBayesModel(5, 0, newick); fg0.setCondProbs(synth.random_pwm(5, 0, 3))
    fg1 = PhyloBayesModel(4, 1, newick); fg1.setCondProbs(synth.random_pwm(4, 1, 4))
    fg0.setBranchLength(2, 3, 0.7777)            # one position with its own branch length (edge learning per position)
    bg = PhyloBackground(0, newick); bg.setCondProb(np.array([[0.3, 0.2, 0.1, 0.4]]), 0)
    return newick, fg0, fg1, bg


def test_phylo_models_round_trip():
    newick, fg0, fg1, bg = _models()
    for m in (fg0, fg1):
        text = xmlio.phylo_bayes_model_to_xml(m)
        r = xmlio.read_phylo_bayes_model(text)
        assert (r.length, r.order) == (m.length, m.order)
        for a, b in zip(m.getCondProbs(), r.getCondProbs()):
            assert np.array_equal(np.asarray(a), np.asarray(b))                 # Double.toString round-trips exactly
        assert np.array_equal(r._branch, np.round(m._branch * 1000.0) / 1000.0)  # VirtualTree.java:223-224
        assert (r.tree.parent == m.tree.parent).all() and r.tree.labels == m.tree.labels
        assert xmlio.phylo_bayes_model_to_xml(r) == xmlio.phylo_bayes_model_to_xml(xmlio.read_phylo_bayes_model(xmlio.phylo_bayes_model_to_xml(r)))
    assert fg0._branch[2, 3] == 0.7777 and xmlio.read_phylo_bayes_model(xmlio.phylo_bayes_model_to_xml(fg0))._branch[2, 3] == 0.778
    t1 = xmlio.parse(xmlio.phylo_bayes_model_to_xml(fg1))
    ct = xmlio.value(t1.child("_bnh").child("connectionTable"))
    assert ct.shape == (4, 4) and ct.sum() == 3 and ct[0, 1] == ct[1, 2] == ct[2, 3] == 1
    assert xmlio.value(t1.child("_bnh").child("VirtualTree1").child("evolModel").child("dimension")) == 4
    assert xmlio.value(t1.child("_bnh").child("VirtualTree0").child("evolModel").child("dimension")) == 1
    rb = xmlio.read_phylo_background(xmlio.phylo_background_to_xml(bg))
    assert rb.order == 0 and np.array_equal(rb.getCondProb(0), bg.getCondProb(0))
    # the stored alphabet is the five-symbol DNA alphabet with the gap symbol
    al = t1.child("alphabets")
    assert al.child("className").text == "de.jstacs.data.AlphabetContainer"
    sym = xmlio.value(al.find("symbols"))
    assert sym == ["A", "C", "G", "T", "-"]


def _mixture_xml(inner, cls, tag, weights, models):
    """jstacs AbstractMixtureModel.toXML: the fields a loader needs, in the reference's order, framed by the class's simple name."""
    x = append_object(JValue("java.lang.Integer", 0), "length")
    x += append_object(JValue("java.lang.Integer", 2), "dimension")
    x += append_object(JValue("java.lang.Integer", 1), "starts")
    x += append_object(JValue("java.lang.Boolean", True), "estimateComponentProbs")
    x += append_object(JValue("[D", [1.0, 1.0]), "componentHyperParams")
    x += append_object(JValue("[Lde.jstacs.models.Model;", models), "models")
    x += append_object(JValue("[Z", [True, False]), "optimizeModel")
    x += append_object(JValue("java.lang.Boolean", True), "algorithmHasBeenRun")
    x += append_object(JValue("[D", weights), "weights")
    x += append_object(JValue("java.lang.Double", 1.0), "alpha")
    x += append_object(None, "best")
    return add_tags(x + inner, tag)


def test_read_single_hidden_motif_mixture_file():
    from phyfoo_b200.models import StrandModel
    newick, fg0, fg1, bg = _models()
    motif = JValue("models.PhyloBayesModel", xmlio.phylo_bayes_model_to_xml(fg0))
    flank = JValue("models.PhyloBackground", xmlio.phylo_background_to_xml(bg))
    plain = _mixture_xml("", "SingleHiddenMotifMixture", "SingleHiddenMotifMixture", [0.7, 0.3], [motif, flank])
    shm = xmlio.read_shm(plain)
    assert shm.strand is None and np.allclose(shm.weights, [0.7, 0.3]) and shm.estimateComponentProbs
    assert np.array_equal(np.asarray(shm.motif.getCondProbs()), np.asarray(fg0.getCondProbs()))
    assert np.array_equal(shm.flanking.getCondProb(0), bg.getCondProb(0))
    # the motif wrapped in a StrandModel, whose own <weights> come before the mixture's in the text
    strand = JValue("de.jstacs.models.mixture.StrandModel",
                    _mixture_xml("", "StrandModel", "StrandModel", [0.55, 0.45], [motif]))
    wrapped = _mixture_xml("", "SingleHiddenMotifMixture", "SingleHiddenMotifMixture", [0.9, 0.1], [strand, flank])
    shm2 = xmlio.read_shm(wrapped)
    assert isinstance(shm2.strand, StrandModel) and np.allclose(shm2.strand.weights, [0.55, 0.45])
    assert np.allclose(shm2.weights, [0.9, 0.1]) and shm2.motif.length == 5
    with pytest.raises(ValueError):
        xmlio.read_shm("<StrandModel>\n</StrandModel>\n")


def test_temperature_other_than_one_is_rejected():
    newick, fg0, fg1, bg = _models()
    text = xmlio.phylo_bayes_model_to_xml(fg0).replace("<TEMPERATURE><className>java.lang.Double</className>1.0",
                                                       "<TEMPERATURE><className>java.lang.Double</className>0.5", 1)
    with pytest.raises(NotImplementedError):
        xmlio.read_phylo_bayes_model(text)


def test_newick_writer_and_parser_round_trip_random_trees():
    """bayesNet/VirtualTree.java:208-234 writes what io/NewickToBayesNet.java:106-147 reads: topology, leaf order and the
    three-decimal branch lengths survive; a second round trip is the identity."""
    from phyfoo_b200 import synth
    from phyfoo_b200.newick import parse_newick
    rng = np.random.default_rng(3)
    for trial in range(40):
        S = int(rng.integers(2, 24))
        text = synth.random_binary_newick(S, int(rng.integers(1, 10 ** 6)), 0.0004, 1.3)
        if trial % 3 == 0:                                   # a multifurcating root as in the star data sets
            text = "(" + ",".join(f"L{k}:{rng.uniform(0.001, 0.9):.4f}" for k in range(S)) + ")"
        t1 = parse_newick(text)
        s1 = xmlio.newick_string(t1, t1.branch)
        t2 = parse_newick(s1)
        assert (t2.parent == t1.parent).all() and t2.labels == t1.labels and (t2.leaf_species == t1.leaf_species).all()
        assert np.array_equal(t2.branch, np.round(t1.branch * 1000.0) / 1000.0)
        assert xmlio.newick_string(t2, t2.branch) == s1
        assert s1.endswith(";") and s1.count("(") == s1.count(")")


def test_mixture_writer_round_trip_and_layout():
    """SingleHiddenMotifMixture / StrandModel files as the library writes them (xmlio.shm_to_xml): the field order of jstacs
    AbstractMixtureModel.toXML (:1628-1678) with the loader's legacy <epsilon> termination condition (:1713-1714), the position
    prior of HiddenMotifMixture.java:208-213; reading them back gives the same models and weights, writing again the same text."""
    from phyfoo_b200.models import SingleHiddenMotifMixture, StrandModel
    newick, fg0, fg1, bg = _models()
    for motif in (fg0, StrandModel(fg0, 0.55), fg1):
        shm = SingleHiddenMotifMixture(motif, bg, zoops=0.9)
        text = shm.toXML(best=-1234.5)
        back = SingleHiddenMotifMixture.fromXML(text)
        assert np.allclose(back.weights, [0.9, 0.1]) and back.estimateComponentProbs is False
        assert (back.strand is None) == (shm.strand is None)
        if shm.strand is not None:
            assert np.allclose(back.strand.weights, [0.55, 0.45]) and back.strand.estimateComponentProbs
        assert xmlio.phylo_bayes_model_to_xml(back.motif) == xmlio.phylo_bayes_model_to_xml(xmlio.read_phylo_bayes_model(xmlio.phylo_bayes_model_to_xml(shm.motif)))
        assert back.toXML(best=-1234.5) == xmlio.shm_to_xml(xmlio.read_shm(back.toXML(best=-1234.5)), best=-1234.5)
        root = xmlio.parse(text).child("SingleHiddenMotifMixture")
        tags = [c.tag for c in root.children if c.tag != "className"]
        assert tags == ["length", "dimension", "starts", "estimateComponentProbs", "componentHyperParams", "models", "optimizeModel",
                        "algorithmHasBeenRun", "weights", "algorithm", "alpha", "epsilon", "parametrization", "best", "posPrior"]
        assert root.child("algorithm").child("className").text == "de.jstacs.models.mixture.AbstractMixtureModel$Algorithm"
        assert root.child("algorithm").child("name").text == "EM" and root.child("parametrization").child("name").text == "LAMBDA"
        assert xmlio.value(root.child("optimizeModel")) == [True, False] and xmlio.value(root.child("best")) == -1234.5
        prior = root.child("posPrior")
        assert prior.child("className").text.endswith("positionprior.UniformPositionPrior")
        assert xmlio.value(prior.child("UniformPositionPrior").child("motifLength")) == shm.motif.getLength()
        assert text.startswith("<SingleHiddenMotifMixture>\n<length><className>java.lang.Integer</className>0</length>")
    # estimated component weights (algorithm.zoops outside (0, 1], models/ModelUtil.java:122-133)
    shm = SingleHiddenMotifMixture(fg0, bg)
    shm.weights = np.array([0.25, 0.75])
    back = xmlio.read_shm(shm.toXML())
    assert back.estimateComponentProbs and np.allclose(back.weights, [0.25, 0.75])
