#!/usr/bin/env python
"""Writes tests/golden/nltk_published_vectors.json: input/output pairs that NLTK ITSELF publishes for the functions behind
the reference's `--constrain` filter (/root/reference/utils_attacks.py:110-143 -> nltk.word_tokenize, i.e. Punkt's
sent_tokenize followed by NLTKWordTokenizer.tokenize).

NLTK is not installed in this image and cannot be fetched (no network), so these vectors could NOT be produced by running
NLTK here. They are the doctest / unit-test examples of the NLTK 3.8.1 sources, transcribed by hand with the file and the
docstring or test they come from; none of the expected outputs was produced by oracle/nltk_restate.py or by the CUDA
kernel - that is the point of the file. `tests/tools/validate_constrain.py --published` re-checks every vector against a real
NLTK install when one is reachable and reports any transcription error.

kind:
  treebank       NLTKWordTokenizer().tokenize(text) / TreebankWordTokenizer().tokenize(text)   (no sentence splitting)
  word_tokenize  nltk.word_tokenize(text)                                                       (Punkt + Treebank)
  sent_tokenize  nltk.sent_tokenize(text)                                                       (Punkt, english.pickle)
"""
import json
import os

MUFFINS = "Good muffins cost $3.88\nin New York.  Please buy me\ntwo of them.\nThanks."
MUFFINS2 = "Good muffins cost $3.88\nin New York.  Please buy me\ntwo of them.\n\nThanks."

V = []


def add(kind, source, text, expected, **kw):
    V.append(dict(kind=kind, source=source, text=text, expected=expected, **kw))


# ---- nltk/tokenize/treebank.py :: TreebankWordTokenizer (class docstring), same examples in destructive.py ----------------
add("treebank", "nltk/tokenize/treebank.py TreebankWordTokenizer docstring; nltk/tokenize/destructive.py NLTKWordTokenizer docstring", MUFFINS,
    ["Good", "muffins", "cost", "$", "3.88", "in", "New", "York.", "Please", "buy", "me", "two", "of", "them.", "Thanks", "."])
add("treebank", "same docstring", "They'll save and invest more.", ["They", "'ll", "save", "and", "invest", "more", "."])
add("treebank", "same docstring", "hi, my name can't hello,", ["hi", ",", "my", "name", "ca", "n't", "hello", ","])
add("treebank", "nltk/tokenize/destructive.py NLTKWordTokenizer.tokenize docstring (convert_parentheses=False branch)",
    "Good muffins cost $3.88 (roughly 3,36 euros)\nin New York.  Please buy me\ntwo of them.\nThanks.",
    ["Good", "muffins", "cost", "$", "3.88", "(", "roughly", "3,36", "euros", ")", "in", "New", "York.", "Please", "buy", "me",
     "two", "of", "them.", "Thanks", "."])
add("treebank", "nltk/tokenize/destructive.py NLTKWordTokenizer.span_tokenize docstring (the token list its spans are checked against)",
    "Good muffins cost $3.88\nin New (York).  Please (buy) me\ntwo of them.\n(Thanks).",
    ["Good", "muffins", "cost", "$", "3.88", "in", "New", "(", "York", ")", ".", "Please", "(", "buy", ")", "me", "two", "of",
     "them.", "(", "Thanks", ")", "."])

# ---- nltk/tokenize/__init__.py module docstring ------------------------------------------------------------------------------
add("word_tokenize", "nltk/tokenize/__init__.py module docstring", MUFFINS2,
    ["Good", "muffins", "cost", "$", "3.88", "in", "New", "York", ".", "Please", "buy", "me", "two", "of", "them", ".", "Thanks", "."])
add("sent_tokenize", "nltk/tokenize/__init__.py module docstring", MUFFINS2,
    ["Good muffins cost $3.88\nin New York.", "Please buy me\ntwo of them.", "Thanks."])

# ---- nltk/test/tokenize.doctest :: "Regression Tests: NLTKWordTokenizer" -----------------------------------------------------
DT = "nltk/test/tokenize.doctest, Regression Tests: NLTKWordTokenizer"
add("word_tokenize", DT, "On a $50,000 mortgage of 30 years at 8 percent, the monthly payment would be $366.88.",
    ["On", "a", "$", "50,000", "mortgage", "of", "30", "years", "at", "8", "percent", ",", "the", "monthly", "payment", "would",
     "be", "$", "366.88", "."])
add("word_tokenize", DT, "\"We beat some pretty good teams to get here,\" Slocum said.",
    ["``", "We", "beat", "some", "pretty", "good", "teams", "to", "get", "here", ",", "''", "Slocum", "said", "."])
add("word_tokenize", DT, "Well, we couldn't have this predictable, cliche-ridden, \"Touched by an Angel\" (a show creator John Masius "
    "worked on) wanna-be if she didn't.",
    ["Well", ",", "we", "could", "n't", "have", "this", "predictable", ",", "cliche-ridden", ",", "``", "Touched", "by", "an",
     "Angel", "''", "(", "a", "show", "creator", "John", "Masius", "worked", "on", ")", "wanna-be", "if", "she", "did", "n't", "."])
add("word_tokenize", DT, "I cannot cannot work under these conditions!",
    ["I", "can", "not", "can", "not", "work", "under", "these", "conditions", "!"])
add("word_tokenize", DT, "The company spent $30,000,000 last year.", ["The", "company", "spent", "$", "30,000,000", "last", "year", "."])
add("word_tokenize", DT, "The company spent 40.75% of its income last year.",
    ["The", "company", "spent", "40.75", "%", "of", "its", "income", "last", "year", "."])
add("word_tokenize", DT, "He arrived at 3:00 pm.", ["He", "arrived", "at", "3:00", "pm", "."])
add("word_tokenize", DT, "I bought these items: books, pencils, and pens.",
    ["I", "bought", "these", "items", ":", "books", ",", "pencils", ",", "and", "pens", "."])
add("word_tokenize", DT, "Though there were 150, 100 of them were old.",
    ["Though", "there", "were", "150", ",", "100", "of", "them", "were", "old", "."])
add("word_tokenize", DT, "There were 300,000, but that wasn't enough.",
    ["There", "were", "300,000", ",", "but", "that", "was", "n't", "enough", "."])
add("word_tokenize", DT, "It's more'n enough.", ["It", "'s", "more", "'n", "enough", "."])
add("word_tokenize", "nltk/test/tokenize.doctest, single quotes (issue #2126); nltk/test/unit/test_tokenize.py::test_word_tokenize",
    "The 'v', I've been fooled but I'll seek revenge.",
    ["The", "'", "v", "'", ",", "I", "'ve", "been", "fooled", "but", "I", "'ll", "seek", "revenge", "."])
add("word_tokenize", "same", "'v' 're'", ["'", "v", "'", "'re", "'"])

# ---- nltk/test/unit/test_tokenize.py -------------------------------------------------------------------------------------------
UT = "nltk/test/unit/test_tokenize.py"
add("word_tokenize", UT + "::test_pad_asterisk", "This is a, *weird sentence with *asterisks in it.",
    ["This", "is", "a", ",", "*", "weird", "sentence", "with", "*", "asterisks", "in", "it", "."])
add("word_tokenize", UT + "::test_pad_dotdot", "Why did dotdot.. not get tokenized but dotdotdot... did? How about manydots.....",
    ["Why", "did", "dotdot", "..", "not", "get", "tokenized", "but", "dotdotdot", "...", "did", "?", "How", "about", "manydots", "....."])
ST = UT + "::test_sent_tokenize (parametrized)"
add("sent_tokenize", ST, "this is a test. . new sentence.", ["this is a test.", ".", "new sentence."])
add("sent_tokenize", ST, "This. . . That", ["This.", ".", ".", "That"])
add("sent_tokenize", ST, "This..... That", ["This..... That"])
add("sent_tokenize", ST, "This... That", ["This... That"])
add("sent_tokenize", ST, "This.. . That", ["This.. .", "That"])
add("sent_tokenize", ST, "This. .. That", ["This.", ".. That"])
add("sent_tokenize", ST, "This. ,. That", ["This.", ",.", "That"])
add("sent_tokenize", ST, "This!!! That", ["This!!!", "That"])
add("sent_tokenize", ST, "This! That", ["This!", "That"])
add("sent_tokenize", ST, "1. This is R .\n2. This is A .\n3. That's all", ["1.", "This is R .", "2.", "This is A .", "3.", "That's all"],
    note="upper-case sentence starters after the numbers: Punkt's orthographic heuristic; the reference lower-cases its input")
add("sent_tokenize", ST, "Hello.\tThere", ["Hello.", "There"])

# ---- nltk/tokenize/punkt.py module docstring (english.pickle, realign_boundaries=True) -----------------------------------------
PK = "nltk/tokenize/punkt.py module docstring"
add("sent_tokenize", PK, "Punkt knows that the periods in Mr. Smith and Johann S. Bach\ndo not mark sentence boundaries.  And sometimes sentences\n"
    "can start with non-capitalized words.  i is a good variable\nname.",
    ["Punkt knows that the periods in Mr. Smith and Johann S. Bach\ndo not mark sentence boundaries.",
     "And sometimes sentences\ncan start with non-capitalized words.", "i is a good variable\nname."], abbrev=["mr"])
add("sent_tokenize", PK, "(How does it deal with this parenthesis?)  \"It should be part of the\nprevious sentence.\" \"(And the same with this one.)\" "
    "('And this one!')\n\"('(And (this)) '?)\" [(and this. )]",
    ["(How does it deal with this parenthesis?)", "\"It should be part of the\nprevious sentence.\"", "\"(And the same with this one.)\"",
     "('And this one!')", "\"('(And (this)) '?)\"", "[(and this. )]"])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nltk_published_vectors.json")
json.dump(dict(nltk_version="3.8.1 (sources as published; transcribed by hand, NLTK not installable here)", vectors=V), open(out, "w"), indent=1)
print(len(V), "vectors ->", out)
